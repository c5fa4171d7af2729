/* abi_smoke.c -- the drop-in boundary used from plain C (C99), as a cgo / pgrx-sys / JNI binding would see it.
 *
 *   gcc -std=c99 -pedantic -Wall -Werror -Iinclude examples/abi_smoke.c -Lpgvector_rx_b200 -lpgvrx_b200 -o abi_smoke
 *
 * Without a GPU it exercises the host-only entry points and checks that the compute entry points fail with a status
 * code and a message instead of crashing (the library has no CPU fallback).  With a GPU (argv[1] = "gpu") it also
 * builds a small index, scans it, inserts into it and runs the exact flat path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pgvrx_b200.h"

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

static float frand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return (float)((*s >> 8) & 0xffff) / 65536.0f - 0.5f; }

int main(int argc, char **argv) {
    const int want_gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
    CHECK(pgvrx_abi_version() == PGVRX_ABI_VERSION);
    CHECK(pgvrx_row_bytes(PGVRX_ELEM_F32, 128) == 512);
    CHECK(pgvrx_row_bytes(PGVRX_ELEM_F16, 768) == 1536);
    CHECK(pgvrx_row_bytes(PGVRX_ELEM_BIT, 1024) == 128);
    CHECK(pgvrx_max_level(16) > 0);
    CHECK(pgvrx_level_from_uniform(0.5, 16) == 0);

    /* FUNCTION 2 + normalise (vector.rs:106-126): [3, 4] -> [0.6, 0.8], norm 5 */
    {
        float in[2] = {3.0f, 4.0f}, out[2];
        double norm = 0.0;
        CHECK(pgvrx_normalize(PGVRX_ELEM_F32, 2, in, out, &norm) == PGVRX_OK);
        CHECK(norm == 5.0 && out[0] == 0.6f && out[1] == 0.8f);
    }
    /* a relation that is not an HNSW index is refused with the reference's message (scan.rs:134-137) */
    {
        unsigned char *pages = (unsigned char *)calloc(2, 8192);
        pgvrx_pages_info info;
        CHECK(pages != NULL);
        CHECK(pgvrx_pages_inspect(pages, 2, PGVRX_ELEM_F32, &info) == PGVRX_ERR_INVALID);
        CHECK(strstr(pgvrx_last_error(NULL), "hnsw index is not valid") != NULL);
        free(pages);
    }

    /* amgettuple's fan-out (scan.rs:798): element 1 holds TIDs {7, 8, 9}, element 0 holds {5}; -1 is result padding */
    {
        uint64_t tids[2 * PGVRX_HEAPTIDS], out[8];
        uint8_t counts[2] = {1, 3};
        int32_t ids[3] = {1, -1, 0};
        int64_t n_out = 0;
        memset(tids, 0xff, sizeof tids);
        tids[0] = 5; tids[PGVRX_HEAPTIDS] = 7; tids[PGVRX_HEAPTIDS + 1] = 8; tids[PGVRX_HEAPTIDS + 2] = 9;
        CHECK(pgvrx_heap_tid_stream(ids, 3, tids, counts, 2, out, 8, &n_out) == PGVRX_OK);
        CHECK(n_out == 4 && out[0] == 9 && out[1] == 8 && out[2] == 7 && out[3] == 5);
    }

    if (!want_gpu) {
        pgvrx_index_desc d;
        pgvrx_index *ix = NULL;
        int rc;
        memset(&d, 0, sizeof d);
        d.elem_type = PGVRX_ELEM_F32; d.metric = PGVRX_METRIC_L2SQ; d.dim = 8; d.m = 16; d.ef_construction = 64; d.capacity = 16;
        rc = pgvrx_index_create(&d, &ix);
        if (pgvrx_device_count() <= 0) {
            CHECK(rc != PGVRX_OK && ix == NULL);          /* no device: fail loudly, never fall back to the CPU */
            CHECK(strlen(pgvrx_last_error(NULL)) > 0);
        } else if (rc == PGVRX_OK) {
            pgvrx_index_destroy(ix);
        }
        printf("abi_smoke: host-only checks passed\n");
        return 0;
    }

    {
        enum { N = 2000, NI = 500, DIM = 32, NQ = 64, K = 10 };
        pgvrx_index_desc d;
        pgvrx_index *ix = NULL;
        pgvrx_build_stats bs;
        float *rows = (float *)malloc(sizeof(float) * (N + NI) * DIM), *q = (float *)malloc(sizeof(float) * NQ * DIM);
        int32_t *levels = (int32_t *)calloc(N + NI, sizeof(int32_t));
        int32_t *ids = (int32_t *)malloc(sizeof(int32_t) * NQ * K), *fids = (int32_t *)malloc(sizeof(int32_t) * NQ * K);
        float *dist = (float *)malloc(sizeof(float) * NQ * K), *fdist = (float *)malloc(sizeof(float) * NQ * K);
        int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * NQ);
        int64_t n = 0;
        unsigned seed = 7;
        int i, hits = 0;
        CHECK(rows && q && levels && ids && fids && dist && fdist && cnt);
        for (i = 0; i < (N + NI) * DIM; i++) rows[i] = frand(&seed);
        for (i = 0; i < N + NI; i++) levels[i] = pgvrx_level_from_uniform((double)((seed = seed * 1664525u + 1013904223u) >> 8) / 16777216.0 + 1e-9, 16);
        memcpy(q, rows, sizeof(float) * NQ * DIM);         /* query with indexed rows: each must find itself first */
        memset(&d, 0, sizeof d);
        d.elem_type = PGVRX_ELEM_F32; d.metric = PGVRX_METRIC_L2SQ; d.dim = DIM; d.m = 16; d.ef_construction = 64; d.capacity = N + NI;
        CHECK(pgvrx_index_create(&d, &ix) == PGVRX_OK);
        CHECK(pgvrx_index_build(ix, rows, N, levels, NULL, 0, &bs) == PGVRX_OK);
        CHECK(pgvrx_index_insert(ix, rows + (size_t)N * DIM, NI, levels + N, NULL, 0, &bs) == PGVRX_OK);
        CHECK(pgvrx_index_graph_info(ix, &n, NULL, NULL, NULL) == PGVRX_OK && n == N + NI);
        CHECK(pgvrx_search_batch(ix, q, NQ, 40, K, ids, dist, cnt, NULL) == PGVRX_OK);
        CHECK(pgvrx_flat_topk(ix, q, NQ, K, fids, fdist, NULL) == PGVRX_OK);
        for (i = 0; i < NQ; i++) {
            CHECK(cnt[i] == K && ids[i * K] == i && dist[i * K] == 0.0f);
            CHECK(fids[i * K] == i && fdist[i * K] == 0.0f);
            hits += ids[i * K + 1] == fids[i * K + 1];
        }
        CHECK(hits > NQ / 2);                               /* the graph scan agrees with the exact scan most of the time */
        CHECK(pgvrx_search_batch(ix, q, NQ, 0, K, ids, dist, cnt, NULL) == PGVRX_ERR_INVALID);
        CHECK(strstr(pgvrx_last_error(ix), "ef_search") != NULL);
        pgvrx_index_destroy(ix);
        free(rows); free(q); free(levels); free(ids); free(fids); free(dist); free(fdist); free(cnt);
        printf("abi_smoke: gpu checks passed (%d/%d second neighbours agree with the exact scan)\n", hits, NQ);
    }
    return 0;
}
