// iter_host.cu -- C-ABI of the batched iterative index scan (K3): pgvrx_scan_begin / _next / _end stand where
// amrescan, the get_scan_items / resume_scan_items calls inside amgettuple, and amendscan stand in the
// reference (src/index/scan.rs:673-791, 794-886).  Per-query scan state lives in HBM behind the handle.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "index_internal.h"

using namespace pgvrx;

struct pgvrx_scan {
    pgvrx_index *ix = nullptr;
    uint64_t generation = 0;      // ix->generation at pgvrx_scan_begin: the visited bitmaps are sized for that graph
    int64_t nq = 0;
    int ef = 0, mode = 0;
    int64_t max_tuples = 0;
    uint32_t n_words = 0;
    int pool_cap = 0, pool_used = 0;
    int reserve_chunks = 0;
    int64_t active = 0;
    int rounds = 0;
    int R = 0, grid = 0, tail_cap = 0;
    uint32_t stage_stride = 0;
    size_t smem_block = 0;
    DevBuf q, bitmap, state, pool, cmin, cfill, next, pool_next, ids, dists, counts, done, want, status, ctr, vis_up, work;
    std::vector<uint8_t> h_done;
};

namespace {

void scan_release(pgvrx_scan *s) {
    DevBuf *b[] = {&s->q, &s->bitmap, &s->state, &s->pool, &s->cmin, &s->cfill, &s->next, &s->pool_next, &s->ids, &s->dists, &s->counts,
                   &s->done, &s->want, &s->status, &s->ctr, &s->vis_up, &s->work};
    for (DevBuf *x : b) x->release();
}

// make room for one more round of every active query (a round appends at most a few keys per distance evaluation)
int ensure_pool(pgvrx_scan *s, int64_t active) {
    pgvrx_index *ix = s->ix;
    // pool_used / pool_cap / reserve_chunks count EXTENTS (ITER_EXT chunks of ITER_CHUNK keys each)
    const int64_t want = (int64_t)s->pool_used + active * s->reserve_chunks + 8;
    if (want > (int64_t)0x7ffffff0 / ITER_EXT) return pgvrx_fail(ix, PGVRX_ERR_NOMEM, "discarded pool would exceed 2^31 chunks");
    if (want <= s->pool_cap) return PGVRX_OK;
    const int64_t cap = want + want / 4;
    const size_t ext_keys = (size_t)ITER_EXT * ITER_CHUNK;
    CU_TRY(ix, s->pool.ensure_keep((size_t)cap * ext_keys * 8, (size_t)s->pool_used * ext_keys * 8));
    CU_TRY(ix, s->cmin.ensure_keep((size_t)cap * ITER_EXT * 8, (size_t)s->pool_used * ITER_EXT * 8));
    CU_TRY(ix, s->cfill.ensure_keep((size_t)cap * ITER_EXT * 4, (size_t)s->pool_used * ITER_EXT * 4));
    CU_TRY(ix, s->next.ensure_keep((size_t)cap * 4, (size_t)s->pool_used * 4));
    s->pool_cap = (int)cap;
    return PGVRX_OK;
}

}  // namespace

extern "C" {

int pgvrx_scan_begin(pgvrx_index *ix, const void *queries, int64_t nq, int32_t ef_search, int32_t iterative_mode,
                     int64_t max_scan_tuples, pgvrx_scan **out) {
    if (!ix || !out) return PGVRX_ERR_INVALID;
    *out = nullptr;
    if (!ix->loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no graph loaded");
    // GUC range messages as PostgreSQL prints them for the reference (tests/pg_regress/expected/hnsw_vector.out:167-185)
    char msg[128];
    if (ef_search < PGVRX_MIN_EF_SEARCH || ef_search > PGVRX_MAX_EF_SEARCH) {                             // options.rs:157-166
        snprintf(msg, sizeof msg, "%d is outside the valid range for parameter \"hnsw.ef_search\" (1 .. 1000)", ef_search);
        return pgvrx_fail(ix, PGVRX_ERR_INVALID, msg);
    }
    if (iterative_mode < PGVRX_ITER_OFF || iterative_mode > PGVRX_ITER_STRICT) {                          // options.rs:168-175
        snprintf(msg, sizeof msg, "invalid value for parameter \"hnsw.iterative_scan\": \"%d\"", iterative_mode);
        return pgvrx_fail(ix, PGVRX_ERR_INVALID, msg);
    }
    if (max_scan_tuples < 1 || max_scan_tuples > (int64_t)0x7fffffff) {                                   // options.rs:177-186
        snprintf(msg, sizeof msg, "%lld is outside the valid range for parameter \"hnsw.max_scan_tuples\" (1 .. 2147483647)", (long long)max_scan_tuples);
        return pgvrx_fail(ix, PGVRX_ERR_INVALID, msg);
    }
    if (nq < 0 || nq > (int64_t)1 << 24 || (nq > 0 && !queries)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad queries");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    pgvrx_scan *s = new (std::nothrow) pgvrx_scan();
    if (!s) return pgvrx_fail(ix, PGVRX_ERR_NOMEM, "out of host memory");
    s->ix = ix;
    s->generation = ix->generation;
    s->nq = nq;
    s->ef = ef_search;
    s->mode = iterative_mode;
    s->max_tuples = max_scan_tuples;
    s->n_words = (uint32_t)((std::max<int64_t>(ix->n, 1) + 31) / 32);
    {
        // a round appends at most a few keys per distance evaluation; + the partially used extent a query may start with
        const int64_t chunks = ((int64_t)ef_search * 2 * ix->desc.m * 2 + 256 + ITER_CHUNK - 1) / ITER_CHUNK + 2;
        s->reserve_chunks = (int)((chunks + ITER_EXT - 1) / ITER_EXT) + 1;
    }
    s->active = nq;
    s->h_done.assign((size_t)nq, 0);
    // staging geometry: the same rule as the one-shot scan (capi.cu), result list always in shared memory
    s->stage_stride = ix->row_stride + (((ix->row_stride / 16) & 1u) ? 0u : 16u);
    s->R = (int)std::max<uint32_t>(1, std::min<uint32_t>(32, (17u * 1024u) / s->stage_stride));
    // slots for retained boundary ties: integer-valued distances (bit types) tie constantly, float distances only on
    // duplicated rows -- fewer slots leave room for more resident warps
    s->tail_cap = ix->desc.elem_type == PGVRX_ELEM_BIT ? ITER_TAIL_CAP : ITER_TAIL_CAP_FLOAT;
    if (const char *e = getenv("PGVRX_ITER_TAIL")) { int v = atoi(e); if (v >= 16 && v <= 4096) s->tail_cap = v; }
    s->smem_block = iter_warp_smem(s->R, s->stage_stride, ix->row_stride, ef_search, s->tail_cap) * ITER_WARPS;
    auto bail = [&](int rc) { scan_release(s); delete s; return rc; };
    if (s->smem_block > 227 * 1024) return bail(pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "scan configuration exceeds shared memory"));
    int occ = 0;
    cudaError_t e = iter_kernel_occupancy(ix->desc.elem_type, ix->desc.metric, s->smem_block, &occ);
    if (e != cudaSuccess || occ < 1) { cudaGetLastError(); return bail(pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "iterative scan kernel does not fit on an SM")); }
    const int wpb = ITER_WARPS;
    s->grid = (int)std::max<int64_t>(1, std::min<int64_t>((nq + wpb - 1) / wpb, (int64_t)ix->sm_count * occ));

#define SCAN_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            std::string _m = std::string(cudaGetErrorName(_e)) + ": " + cudaGetErrorString(_e);     \
            cudaGetLastError();                                                                     \
            return bail(pgvrx_fail(ix, _e == cudaErrorMemoryAllocation ? PGVRX_ERR_NOMEM : PGVRX_ERR_CUDA, _m)); \
        }                                                                                           \
    } while (0)

    const size_t nqz = (size_t)std::max<int64_t>(nq, 1);
    SCAN_TRY(s->q.ensure(nqz * ix->row_bytes));
    SCAN_TRY(s->bitmap.ensure(nqz * s->n_words * 4));
    SCAN_TRY(s->state.ensure(nqz * sizeof(IterQueryState)));
    SCAN_TRY(s->ids.ensure(nqz * ef_search * 4));
    SCAN_TRY(s->dists.ensure(nqz * ef_search * 4));
    SCAN_TRY(s->counts.ensure(nqz * 4));
    SCAN_TRY(s->done.ensure(nqz));
    SCAN_TRY(s->want.ensure(nqz));
    SCAN_TRY(s->status.ensure(nqz * 4));
    SCAN_TRY(s->ctr.ensure(nqz * 16));
    SCAN_TRY(s->vis_up.ensure((size_t)s->grid * wpb * ITER_UP_ENTRIES * 4));
    SCAN_TRY(s->work.ensure(64));
    SCAN_TRY(s->pool_next.ensure(64));
    if (nq > 0) {
        SCAN_TRY(cudaMemcpyAsync(s->q.p, queries, (size_t)nq * ix->row_bytes, cudaMemcpyHostToDevice, ix->stream));
        SCAN_TRY(cudaMemsetAsync(s->bitmap.p, 0, (size_t)nq * s->n_words * 4, ix->stream));
        std::vector<IterQueryState> init((size_t)nq);
        for (auto &st : init) { memset(&st, 0, sizeof st); st.head = st.tail = -1; }
        SCAN_TRY(cudaMemcpyAsync(s->state.p, init.data(), init.size() * sizeof(IterQueryState), cudaMemcpyHostToDevice, ix->stream));
        SCAN_TRY(cudaMemsetAsync(s->pool_next.p, 0, 64, ix->stream));
        SCAN_TRY(cudaStreamSynchronize(ix->stream));
    }
#undef SCAN_TRY
    *out = s;
    return PGVRX_OK;
}

int pgvrx_scan_next(pgvrx_scan *s, const uint8_t *want_more, int32_t *out_ids, float *out_dists, int32_t *out_counts,
                    uint8_t *out_done, pgvrx_search_stats *stats) {
    return pgvrx_scan_next_rounds(s, want_more, 1, out_ids, out_dists, out_counts, out_done, stats);
}

int pgvrx_scan_next_rounds(pgvrx_scan *s, const uint8_t *want_more, int32_t rounds, int32_t *out_ids, float *out_dists,
                           int32_t *out_counts, uint8_t *out_done, pgvrx_search_stats *stats) {
    if (!s) return PGVRX_ERR_INVALID;
    pgvrx_index *ix = s->ix;
    if (stats) memset(stats, 0, sizeof *stats);
    if (rounds < 1 || rounds > 1024) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "rounds must be between 1 and 1024");
    if (s->nq == 0) return PGVRX_OK;
    if (!out_ids || !out_dists || !out_counts || !out_done) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null output buffer");
    if (!ix->loaded || ix->generation != s->generation)
        return pgvrx_fail(ix, PGVRX_ERR_STATE, "the index changed under an open scan (build, insert, vacuum, load or flags since "
                                               "pgvrx_scan_begin): end this scan and begin a new one");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    cudaStream_t stream = ix->stream;
    const int64_t nq = s->nq;
    int64_t active = 0;
    for (int64_t i = 0; i < nq; i++) active += !s->h_done[(size_t)i] && (!want_more || want_more[i]);
    int rc = ensure_pool(s, active * rounds);
    if (rc) return rc;
    CU_TRY(ix, s->ids.ensure((size_t)nq * rounds * s->ef * 4));
    CU_TRY(ix, s->dists.ensure((size_t)nq * rounds * s->ef * 4));
    CU_TRY(ix, s->counts.ensure((size_t)nq * rounds * 4));
    if (want_more) CU_TRY(ix, cudaMemcpyAsync(s->want.p, want_more, (size_t)nq, cudaMemcpyHostToDevice, stream));
    CU_TRY(ix, cudaMemsetAsync(s->work.p, 0, 64, stream));
    CU_TRY(ix, cudaEventRecord(ix->ev0, stream));

    IterParams p;
    p.queries = (const uint8_t *)s->q.p;
    p.q_stride = (uint32_t)ix->row_bytes;
    p.q_bytes = (uint32_t)ix->row_bytes;
    p.nq = (int)nq;
    p.ef = s->ef;
    p.tail_cap = s->tail_cap;
    p.mode = s->mode;
    p.rounds = rounds;
    p.max_scan_tuples = s->max_tuples;
    p.want_more = want_more ? (const uint8_t *)s->want.p : nullptr;
    p.state = (IterQueryState *)s->state.p;
    p.bitmap = (uint32_t *)s->bitmap.p;
    p.n_words = s->n_words;
    p.pool = (uint64_t *)s->pool.p;
    p.cmin = (uint64_t *)s->cmin.p;
    p.cfill = (int *)s->cfill.p;
    p.next = (int *)s->next.p;
    p.pool_next = (int *)s->pool_next.p;
    p.pool_cap = s->pool_cap;
    p.vis_up = (uint32_t *)s->vis_up.p;
    p.out_ids = (int32_t *)s->ids.p;
    p.out_dists = (float *)s->dists.p;
    p.out_counts = (int32_t *)s->counts.p;
    p.out_done = (uint8_t *)s->done.p;
    p.out_status = (uint32_t *)s->status.p;
    p.out_ctr = (uint32_t *)s->ctr.p;
    p.work_counter = (int *)s->work.p;
    p.stage_stride = s->stage_stride;
    p.R = s->R;
    CU_TRY(ix, launch_iter_scan(ix->desc.elem_type, ix->desc.metric, pgvrx_view_of(ix), p, s->grid, s->smem_block, stream));
    CU_TRY(ix, cudaEventRecord(ix->ev1, stream));
    CU_TRY(ix, cudaMemcpyAsync(out_ids, s->ids.p, (size_t)nq * rounds * s->ef * 4, cudaMemcpyDeviceToHost, stream));
    CU_TRY(ix, cudaMemcpyAsync(out_dists, s->dists.p, (size_t)nq * rounds * s->ef * 4, cudaMemcpyDeviceToHost, stream));
    CU_TRY(ix, cudaMemcpyAsync(out_counts, s->counts.p, (size_t)nq * rounds * 4, cudaMemcpyDeviceToHost, stream));
    CU_TRY(ix, cudaMemcpyAsync(out_done, s->done.p, (size_t)nq, cudaMemcpyDeviceToHost, stream));
    std::vector<uint32_t> st((size_t)nq), ctr((size_t)nq * 4);
    int used = 0;
    CU_TRY(ix, cudaMemcpyAsync(st.data(), s->status.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
    CU_TRY(ix, cudaMemcpyAsync(ctr.data(), s->ctr.p, (size_t)nq * 16, cudaMemcpyDeviceToHost, stream));
    CU_TRY(ix, cudaMemcpyAsync(&used, s->pool_next.p, 4, cudaMemcpyDeviceToHost, stream));
    CU_TRY(ix, cudaStreamSynchronize(stream));
    s->pool_used = std::min(used, s->pool_cap);
    s->rounds += rounds;
    for (int64_t i = 0; i < nq; i++) {
        s->h_done[(size_t)i] = out_done[i];
        if (st[(size_t)i] != 0) {
            char b[200];
            snprintf(b, sizeof b, "query %lld exhausted the iterative scan's scratch (status %u: %s)", (long long)i, st[(size_t)i],
                     (st[(size_t)i] & 4u) ? "discarded pool" : (st[(size_t)i] & 2u) ? "too many elements tie with the furthest result" : "visited table");
            return pgvrx_fail(ix, (st[(size_t)i] & 4u) ? PGVRX_ERR_NOMEM : PGVRX_ERR_UNSUPPORTED, b);
        }
    }
    if (stats) {
        CU_TRY(ix, cudaEventElapsedTime(&stats->kernel_ms, ix->ev0, ix->ev1));
        stats->n_launches = 1;
        for (int64_t i = 0; i < nq; i++) {
            stats->n_dist += ctr[(size_t)i * 4 + 0];
            stats->n_expand0 += ctr[(size_t)i * 4 + 1];
            stats->n_expand_upper += ctr[(size_t)i * 4 + 2];
            stats->n_visited += ctr[(size_t)i * 4 + 3];
        }
    }
    return PGVRX_OK;
}

void pgvrx_scan_end(pgvrx_scan *s) {
    if (!s) return;
    cudaSetDevice(s->ix->desc.device);
    scan_release(s);
    delete s;
}

}  // extern "C"
