// capi.cu -- implementation of include/pgvrx_b200.h: handle, HBM-resident graph, launches.
//
// Nothing here computes a distance or walks a graph on the CPU.  If no CUDA device is usable every
// compute entry point fails with PGVRX_ERR_CUDA: there is no fallback path.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "index_internal.h"

using namespace pgvrx;

namespace {
thread_local std::string g_create_error;
}

namespace pgvrx_host {
void set_create_error(const std::string &msg) { g_create_error = msg; }
}

int pgvrx_fail(pgvrx_index *idx, int code, const std::string &msg) {
    if (idx) idx->err = msg; else g_create_error = msg;
    return code;
}

GraphView pgvrx_view_of(const pgvrx_index *ix) {
    GraphView g;
    g.vectors = (const uint8_t *)ix->vectors.p;
    g.adj0 = (const int32_t *)ix->adj0.p;
    g.adj_up = (const int32_t *)ix->adj_up.p;
    g.upper_base = (const int32_t *)ix->upper_base.p;
    g.level = (const uint8_t *)ix->level.p;
    g.flags = ix->has_flags ? (const uint8_t *)ix->flags.p : nullptr;
    g.n = (int32_t)ix->n;
    g.m = ix->desc.m;
    g.dim = ix->desc.dim;
    g.entry = (int32_t)ix->entry;
    g.entry_level = ix->entry_level;
    g.row_stride = ix->row_stride;
    return g;
}

namespace {

inline int fail(pgvrx_index *idx, int code, const std::string &msg) { return pgvrx_fail(idx, code, msg); }
inline GraphView view_of(const pgvrx_index *ix) { return pgvrx_view_of(ix); }

bool metric_ok(int elem, int metric) {
    if (elem == PGVRX_ELEM_BIT) return metric == PGVRX_METRIC_HAMMING || metric == PGVRX_METRIC_JACCARD;
    if (elem == PGVRX_ELEM_F32 || elem == PGVRX_ELEM_F16)
        return metric == PGVRX_METRIC_L2SQ || metric == PGVRX_METRIC_NEG_IP || metric == PGVRX_METRIC_L1;
    return false;
}

int max_dim(int elem) {
    return elem == PGVRX_ELEM_F32 ? PGVRX_MAX_DIM_VECTOR : elem == PGVRX_ELEM_F16 ? PGVRX_MAX_DIM_HALFVEC
                                                                                 : PGVRX_MAX_DIM_BIT;
}

}  // namespace

uint32_t pgvrx_pow2_at_least(uint64_t x) {
    uint32_t p = 256;
    while (p < x && p < (1u << 30)) p <<= 1;
    return p;
}

// tail_cap < 0: pick the register-resident list when ef is small enough (capacity 32*J - 1 with at
// least 16 retained-tie slots), else the shared-memory list with 32 tie slots
int pgvrx_make_config(pgvrx_index *ix, int ef, int tail_cap, uint32_t vis_entries, int64_t nq, int max_grid,
                      bool for_build, ScanConfig *out) {
    // Bytes of shared memory per warp for staged rows.  Scans: 8.5 KiB = 16 rows of a 128-d fp32 vector, i.e. two gather
    // rounds per 32-entry neighbour list but twice the resident warps (measured on C2, profiles/r2_notes.md: 7.49 M QPS
    // at 16 rows against 7.17 M at 32 and 6.9 M at 11).  The build's searches are latency bound at a few dozen resident
    // warps on the whole device, so they keep all 32 rows in one round.  PGVRX_STAGE_KB / PGVRX_STAGE_ROWS override.
    uint32_t STAGE_BUDGET = for_build ? 17 * 1024 : 8704;
    if (const char *e = getenv("PGVRX_STAGE_KB")) { int kb = atoi(e); if (kb >= 1 && kb <= 48) STAGE_BUDGET = (uint32_t)kb * 1024; }
    ScanConfig c;
    c.L.ef = ef;
    c.L.list_j = 0;
    // an index that carries vacuum flags scans -- and takes inserts -- with the shared-memory list only (one extra kernel
    // per opclass and path, not four)
    const bool flagged = ix->has_flags;
    if (tail_cap < 0) {
        tail_cap = 32;
        if (!flagged) for (int j = 2; j <= 8; j *= 2)
            if (32 * j - 1 - ef >= 16) { c.L.list_j = j; tail_cap = 32 * j - 1 - ef; break; }
    }
    c.L.tail_cap = tail_cap;
    c.L.row_stride = ix->row_stride;
    // LDS.128 by one lane per row is conflict free when consecutive rows start 16 B * odd apart
    c.L.stage_stride = ix->row_stride + (((ix->row_stride / 16) & 1u) ? 0u : 16u);
    c.L.R = (int)std::max<uint32_t>(1, std::min<uint32_t>(32, STAGE_BUDGET / c.L.stage_stride));
    // equal rounds: 32 entries in ceil(32 / R) rounds need no more than ceil(32 / rounds) rows each (21 -> 16)
    if (c.L.R < 32) { const int rounds = (32 + c.L.R - 1) / c.L.R; c.L.R = (32 + rounds - 1) / rounds; }
    if (const char *e = getenv("PGVRX_STAGE_ROWS")) { int r = atoi(e); if (r >= 1 && r <= 32) c.L.R = r; }   // tuning: exact rows per round
    c.L.stage_bytes = (uint32_t)c.L.R * c.L.stage_stride;
    c.L.chunk_bytes = c.L.chunk_stride = 0;
    // Wide rows (fewer than 17 fit): stage 512-byte chunks of 32 rows instead of a few whole rows, so that every lane has
    // a row to accumulate -- a 768-d halfvec row (1536 B) otherwise keeps 11 of the 32 lanes busy for six rounds per
    // 64-entry neighbour list.  PGVRX_NO_CHUNK=1 keeps whole-row staging (A/B measurements).
    if (c.L.stage_stride > 1056 && !getenv("PGVRX_NO_CHUNK")) {
        c.L.chunk_bytes = 512;
        c.L.chunk_stride = 528;                       // 33 x 16 B: consecutive rows start an odd number of 16-B vectors apart
        c.L.stage_bytes = std::max<uint32_t>(32u * c.L.chunk_stride, c.L.stage_stride);   // (one whole row for the entry point)
    }
    if (const char *e = getenv("PGVRX_CHUNK_BYTES")) {       // tuning: force chunked staging with this chunk size
        const int cb = atoi(e);
        if (cb >= 64 && cb % 16 == 0 && (uint32_t)cb < ix->row_stride) {
            c.L.chunk_bytes = (uint32_t)cb;
            c.L.chunk_stride = (uint32_t)cb + ((((uint32_t)cb / 16) & 1u) ? 0u : 16u);
            c.L.stage_bytes = std::max<uint32_t>(32u * c.L.chunk_stride, c.L.stage_stride);
        }
    }
    // The build's searches are latency bound with shared memory to spare (a few dozen resident warps on the whole device):
    // two chunk buffers, the next chunk in flight while this one is accumulated (hnsw_device.cuh eval_neighbours).
    // PGVRX_CHUNK_DB=0/1 pins the choice for either kind of search.
    if (c.L.chunk_bytes != 0) {
        bool db = for_build;
        if (const char *e = getenv("PGVRX_CHUNK_DB")) db = atoi(e) != 0;
        if (db) c.L.stage_bytes = std::max<uint32_t>(64u * c.L.chunk_stride, c.L.stage_stride);
    }
    c.vis_entries = vis_entries;
    // scans run one warp (= one query in flight) per CTA: a CTA's slot is handed to the next batch as soon as that warp
    // has finished its last query instead of waiting for three siblings (measured 6.86 / 6.71 / 6.58 M QPS for 1 / 2 / 4
    // warps per CTA, profiles/r1_notes.md); PGVRX_SCAN_WARPS overrides
    // (the build's latency-bound rounds measured the same with 1 or 4 warps per CTA: 12.4 s of kernels for 300k rows)
    int wpb = for_build ? SEARCH_THREADS / 32 : 1;
    if (const char *e = getenv(for_build ? "PGVRX_BUILD_WARPS" : "PGVRX_SCAN_WARPS")) { int w = atoi(e); if (w >= 1 && w <= SEARCH_THREADS / 32) wpb = w; }
    if (tail_cap >= 2048) wpb = 1;   // retry / escalation geometry: the tie tail alone is 16 KiB or more per warp
    c.wpb = wpb;
    c.smem_block = search_warp_smem(c.L) * wpb;
    if (c.smem_block > 227 * 1024) return fail(ix, PGVRX_ERR_UNSUPPORTED, "search configuration exceeds shared memory");
    int occ = 0;
    if (for_build) CU_TRY(ix, build_kernel_occupancy(ix->desc.elem_type, ix->desc.metric, c.L.list_j, flagged, c.smem_block, wpb * 32, &occ));
    else CU_TRY(ix, scan_kernel_occupancy(ix->desc.elem_type, ix->desc.metric, c.L.list_j, flagged, c.smem_block, wpb * 32, &occ));
    if (occ < 1) return fail(ix, PGVRX_ERR_UNSUPPORTED, "search kernel does not fit on an SM");
    int64_t want = (nq + wpb - 1) / wpb;
    int64_t cap = (int64_t)ix->sm_count * occ;
    if (max_grid > 0) cap = std::min<int64_t>(cap, max_grid);
    c.grid = (int)std::max<int64_t>(1, std::min(want, cap));
    *out = c;
    return PGVRX_OK;
}

namespace {

int search_core(pgvrx_index *ix, const void *d_queries, uint32_t q_stride, int64_t nq, int ef, int k,
                int32_t *d_ids, float *d_dists, int32_t *d_counts, pgvrx_search_stats *stats,
                cudaStream_t stream, int64_t *n_failed /* NULL: asynchronous, per-query status stays on the device */,
                int64_t *first_failed) {
    if (!ix->loaded) return fail(ix, PGVRX_ERR_STATE, "no graph loaded");
    if (ef < PGVRX_MIN_EF_SEARCH || ef > PGVRX_MAX_EF_SEARCH) {                         // options.rs:157-166, hnsw_vector.out:167-169
        char b[112];
        snprintf(b, sizeof b, "%d is outside the valid range for parameter \"hnsw.ef_search\" (1 .. 1000)", ef);
        return fail(ix, PGVRX_ERR_INVALID, b);
    }
    if (k < 1 || nq < 0 || nq > (int64_t)1 << 30) return fail(ix, PGVRX_ERR_INVALID, "bad k or nq");
    if (stats) memset(stats, 0, sizeof *stats);
    if (nq == 0) return PGVRX_OK;
    CU_TRY(ix, cudaSetDevice(ix->desc.device));

    const int m = ix->desc.m;
    ScanConfig main_cfg, retry_cfg;
    uint32_t vis_main = pgvrx_pow2_at_least(std::max<uint64_t>(1024, (uint64_t)ef * 2 * m * 4));   // <= ~25 % load
    int rc = pgvrx_make_config(ix, ef, -1, vis_main, nq, 0, false, &main_cfg);
    if (rc) return rc;
    uint32_t vis_retry = pgvrx_pow2_at_least(std::min<uint64_t>((uint64_t)vis_main * 16, std::max<uint64_t>(1024, (uint64_t)ix->n * 2)));
    rc = pgvrx_make_config(ix, ef, 2048, vis_retry, nq, ix->sm_count, false, &retry_cfg);
    if (rc) return rc;

    const int wpb = main_cfg.wpb;
    // scratch set of this stream (see index_internal.h)
    pgvrx_index::ScanScratch *sc = nullptr;
    for (auto &cand : ix->scan_scratch) if (cand.used && cand.owner == stream) { sc = &cand; break; }
    if (!sc) for (auto &cand : ix->scan_scratch) if (!cand.used) { sc = &cand; break; }
    if (!sc) {
        sc = &ix->scan_scratch[ix->scan_scratch_rr];
        ix->scan_scratch_rr = (ix->scan_scratch_rr + 1) % pgvrx_index::N_SCAN_SCRATCH;
        CU_TRY(ix, cudaStreamSynchronize(sc->owner));      // its previous user must be done before another stream touches it
    }
    sc->used = true;
    sc->owner = stream;
    CU_TRY(ix, sc->vis.ensure((size_t)main_cfg.grid * wpb * main_cfg.vis_entries * 4));
    CU_TRY(ix, sc->vis_retry.ensure((size_t)retry_cfg.grid * retry_cfg.wpb * retry_cfg.vis_entries * 4));
    CU_TRY(ix, sc->status.ensure((size_t)nq * 4));
    CU_TRY(ix, sc->qmap.ensure((size_t)nq * 4));
    CU_TRY(ix, sc->work.ensure(64));
    if (stats) CU_TRY(ix, sc->ctr.ensure((size_t)nq * 16));

    int *work = (int *)sc->work.p;   // [0] main counter, [1] retry counter, [2] n_failed
    CU_TRY(ix, cudaMemsetAsync(work, 0, 64, stream));
    if (stats) CU_TRY(ix, cudaEventRecord(ix->ev0, stream));

    GraphView g = view_of(ix);
    SearchParams p;
    p.queries = (const uint8_t *)d_queries;
    p.q_stride = q_stride;
    p.q_bytes = (uint32_t)ix->row_bytes;
    p.qmap = nullptr;
    p.n_work = nullptr;
    p.nq = (int)nq;
    p.ef = ef;
    p.k = k;
    p.tail_cap = main_cfg.L.tail_cap;
    p.out_ids = d_ids;
    p.out_dists = d_dists;
    p.out_counts = d_counts;
    p.out_status = (uint32_t *)sc->status.p;
    p.out_ctr = stats ? (uint32_t *)sc->ctr.p : nullptr;
    p.out_phase = nullptr;
#ifdef PGVRX_PHASE_TIMING
    if (stats) {
        CU_TRY(ix, sc->phase.ensure((size_t)nq * 32));
        CU_TRY(ix, cudaMemsetAsync(sc->phase.p, 0, (size_t)nq * 32, stream));
        p.out_phase = (uint32_t *)sc->phase.p;
    }
#endif
    p.vis_base = (uint32_t *)sc->vis.p;
    p.vis_entries = main_cfg.vis_entries;
    p.up_entries = std::min<uint32_t>(main_cfg.vis_entries, 1024u);   // cleared once per upper layer: keep it small
    p.work_counter = work + 0;
    p.stage_stride = main_cfg.L.stage_stride;
    p.R = main_cfg.L.R;
    p.stage_bytes = main_cfg.L.stage_bytes;
    p.chunk_bytes = main_cfg.L.chunk_bytes;
    p.chunk_stride = main_cfg.L.chunk_stride;
    // Speculative order (fetch every neighbour's row while the visited probes are in flight) when a list's rows go in
    // one round; probe-first (fetch only the rows that will be evaluated: 17 % less L2 -> SM traffic) when they take
    // two or more -- measured on C2: 32 rows/round 7.17 M QPS speculative vs 6.86 M probe-first, 16 rows/round 7.49 M vs
    // 7.68 M (profiles/r2_notes.md).  In between (2, the default for several rounds of whole rows): the FIRST round's rows
    // are requested before the probes are consumed, so the probe's L2 round trip -- the top stall of the probe-first
    // order, 10 % of the samples -- overlaps that gather, and the later rounds fetch only what is evaluated (+0.4 .. 3.6 %
    // over probe-first in back-to-back runs).  PGVRX_SPEC_GATHER=0/1/2 pins the order.
    p.spec = main_cfg.L.chunk_bytes != 0 ? 0 : (main_cfg.L.R >= 32 ? 1 : 2);
    if (getenv("PGVRX_NO_SPEC_GATHER")) p.spec = 0;
    if (const char *e = getenv("PGVRX_SPEC_GATHER")) { const int v = atoi(e); p.spec = v == 2 ? 2 : (v != 0); }   // 2: round 0 speculative, later rounds probe-first
    CU_TRY(ix, launch_scan_kernel(ix->desc.elem_type, ix->desc.metric, main_cfg.L.list_j, g, p, main_cfg.grid, main_cfg.smem_block, main_cfg.wpb * 32, stream));

    // queries that overflowed their visited table or tie tail run again with 16x the scratch
    CU_TRY(ix, launch_collect_failed(p.out_status, (int)nq, (int32_t *)sc->qmap.p, work + 2, stream));
    SearchParams r = p;
    r.qmap = (const int32_t *)sc->qmap.p;
    r.n_work = work + 2;
    r.tail_cap = retry_cfg.L.tail_cap;
    r.vis_base = (uint32_t *)sc->vis_retry.p;
    r.vis_entries = retry_cfg.vis_entries;
    r.up_entries = retry_cfg.vis_entries;   // the retry's greedy descent gets the whole (16x) table as well
    r.work_counter = work + 1;
    r.stage_stride = retry_cfg.L.stage_stride;
    r.R = retry_cfg.L.R;
    r.stage_bytes = retry_cfg.L.stage_bytes;
    r.chunk_bytes = retry_cfg.L.chunk_bytes;
    r.chunk_stride = retry_cfg.L.chunk_stride;
    CU_TRY(ix, launch_scan_kernel(ix->desc.elem_type, ix->desc.metric, retry_cfg.L.list_j, g, r, retry_cfg.grid, retry_cfg.smem_block, retry_cfg.wpb * 32, stream));

    if (stats) {
        CU_TRY(ix, cudaEventRecord(ix->ev1, stream));
        CU_TRY(ix, cudaStreamSynchronize(stream));
        CU_TRY(ix, cudaEventElapsedTime(&stats->kernel_ms, ix->ev0, ix->ev1));
        stats->n_launches = 3;
        std::vector<uint32_t> h((size_t)nq * 4);
        CU_TRY(ix, cudaMemcpy(h.data(), sc->ctr.p, (size_t)nq * 16, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < nq; i++) {
            stats->n_dist += h[(size_t)i * 4 + 0];
            stats->n_expand0 += h[(size_t)i * 4 + 1];
            stats->n_expand_upper += h[(size_t)i * 4 + 2];
            stats->n_visited += h[(size_t)i * 4 + 3];
        }
#ifdef PGVRX_PHASE_TIMING
        {
            std::vector<uint32_t> ph((size_t)nq * 8);
            CU_TRY(ix, cudaMemcpy(ph.data(), sc->phase.p, (size_t)nq * 32, cudaMemcpyDeviceToHost));
            double sum[8] = {0};
            for (int64_t i = 0; i < nq; i++) for (int t = 0; t < 8; t++) sum[t] += ph[(size_t)i * 8 + t];
            fprintf(stderr, "[pgvrx phase cycles/query] pop+adj %.0f  visited %.0f  tma %.0f  distance %.0f  accept %.0f  | total %.0f"
                            "  (per expansion: %.0f %.0f %.0f %.0f %.0f)\n",
                    sum[0] / nq, sum[1] / nq, sum[2] / nq, sum[3] / nq, sum[4] / nq, sum[5] / nq,
                    sum[0] / (stats->n_expand0 + stats->n_expand_upper), sum[1] / (stats->n_expand0 + stats->n_expand_upper),
                    sum[2] / (stats->n_expand0 + stats->n_expand_upper), sum[3] / (stats->n_expand0 + stats->n_expand_upper),
                    sum[4] / (stats->n_expand0 + stats->n_expand_upper));
        }
#endif
        int nf = 0;
        CU_TRY(ix, cudaMemcpy(&nf, work + 2, 4, cudaMemcpyDeviceToHost));
        stats->n_retried = (uint64_t)nf;
    }
    if (n_failed) {
        // per-query outcome: a query that exhausted its scratch even in the retry pass has out_counts = -1; the others
        // are complete.  The caller reports the failure after it has handed out the results of the rest.
        CU_TRY(ix, cudaStreamSynchronize(stream));
        std::vector<uint32_t> st((size_t)nq);
        CU_TRY(ix, cudaMemcpy(st.data(), sc->status.p, (size_t)nq * 4, cudaMemcpyDeviceToHost));
        *n_failed = 0;
        for (int64_t i = 0; i < nq; i++)
            if (st[(size_t)i] != 0) {
                if (*n_failed == 0 && first_failed) *first_failed = i;
                (*n_failed)++;
            }
    }
    return PGVRX_OK;
}

}  // namespace

extern "C" {

int pgvrx_abi_version(void) { return PGVRX_ABI_VERSION; }

const char *pgvrx_last_error(const pgvrx_index *idx) { return idx ? idx->err.c_str() : g_create_error.c_str(); }

int pgvrx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

size_t pgvrx_row_bytes(int32_t elem_type, int32_t dim) {
    if (dim <= 0) return 0;
    switch (elem_type) {
        case PGVRX_ELEM_F32: return (size_t)dim * 4;
        case PGVRX_ELEM_F16: return (size_t)dim * 2;
        case PGVRX_ELEM_BIT: return ((size_t)dim + 7) / 8;
    }
    return 0;
}

// types/hnsw.rs:337-349 with PostgreSQL's 8 KiB page: (8192 - 24 - 8 - 4 - 4) / 6 / m - 2, at most 255
int32_t pgvrx_max_level(int32_t m) {
    if (m < 1) return 0;
    int v = (8192 - 24 - 8 - 4 - 4) / 6 / m - 2;
    return v < 255 ? v : 255;
}

// build.rs:373-377
int32_t pgvrx_level_from_uniform(double r, int32_t m) {
    if (!(r > std::numeric_limits<double>::min())) r = std::numeric_limits<double>::min();
    // the reference multiplies by ml = 1/ln(m) (hnsw_constants.rs:132-134); keep that exact form
    double lv = std::floor(-std::log(r) * (1.0 / std::log((double)m)));
    int cap = pgvrx_max_level(m);
    return lv < (double)cap ? (int32_t)lv : cap;
}

// vector.rs:106-126 / halfvec.rs:204-233.  f32 -> half uses IEEE round-to-nearest-even
// (__float2half_rn) except that the reference flushes |x| < 2^-24 to zero where IEEE would round
// (2^-25, 2^-24) up to the smallest subnormal (halfvec.rs:111-113).
namespace {
// one row; 0 = ok, 1 = halfvec overflow, 2 = the type has no norm.  No error state: also runs on worker threads.
int normalize_row(int32_t elem_type, int32_t dim, const void *in, void *out, double *norm_out) {
    double norm = 0.0;
    if (elem_type == PGVRX_ELEM_F32) {
        const float *a = (const float *)in;
        float *r = (float *)out;
        for (int i = 0; i < dim; i++) norm += (double)a[i] * (double)a[i];
        norm = std::sqrt(norm);
        for (int i = 0; i < dim; i++) r[i] = norm > 0.0 ? (float)((double)a[i] / norm) : 0.0f;
    } else if (elem_type == PGVRX_ELEM_F16) {
        const __half *a = (const __half *)in;
        __half *r = (__half *)out;
        for (int i = 0; i < dim; i++) { double v = (double)__half2float(a[i]); norm += v * v; }
        norm = std::sqrt(norm);
        bool overflow = false;
        for (int i = 0; i < dim; i++) {
            float v = norm > 0.0 ? (float)((double)__half2float(a[i]) / norm) : 0.0f;
            __half h = std::fabs(v) < 5.9604644775390625e-08f ? __float2half_rn(std::signbit(v) ? -0.0f : 0.0f)
                                                              : __float2half_rn(v);
            if (std::isinf(__half2float(h))) overflow = true;
            r[i] = h;
        }
        if (overflow) return 1;
    } else {
        return 2;
    }
    if (norm_out) *norm_out = norm;
    return 0;
}
int normalize_status(int code) {
    if (code == 1) return fail(nullptr, PGVRX_ERR_INVALID, "value out of range: overflow");   // halfvec.rs:225-230
    if (code == 2) return fail(nullptr, PGVRX_ERR_INVALID, "type has no norm support function");
    return PGVRX_OK;
}
}  // namespace

int pgvrx_normalize(int32_t elem_type, int32_t dim, const void *in, void *out, double *norm_out) {
    if (dim <= 0 || !in || !out) return fail(nullptr, PGVRX_ERR_INVALID, "bad normalize arguments");
    return normalize_status(normalize_row(elem_type, dim, in, out, norm_out));
}

// n rows at once (row i at in + i * row_bytes); norms (optional) receives every row's norm.  Rows with norm 0 come out as
// zero rows: the caller drops them from a build (build.rs:433-435) -- queries are normalised the same way (scan.rs:749-754).
// Rows are independent: large batches (a CREATE INDEX worth of halfvec rows is seconds of software half conversions)
// are split over host threads; the status is that of the first failing row, as in a sequential pass.
int pgvrx_normalize_rows(int32_t elem_type, int32_t dim, const void *in, void *out, int64_t n, double *norms) {
    if (n < 0 || (n > 0 && (!in || !out))) return fail(nullptr, PGVRX_ERR_INVALID, "bad normalize arguments");
    const size_t rb = pgvrx_row_bytes(elem_type, dim);
    if (rb == 0) return fail(nullptr, PGVRX_ERR_INVALID, "bad normalize arguments");
    const int64_t work = n * (int64_t)dim;
    int nt = (int)std::min<int64_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())), work / (1 << 20));
    if (n > 0 && (const uint8_t *)out < (const uint8_t *)in + (size_t)n * rb && (const uint8_t *)in < (const uint8_t *)out + (size_t)n * rb && out != in)
        nt = 1;                                          // partially overlapping buffers: keep the sequential order
    nt = std::max(nt, 1);
    std::vector<int64_t> first_bad((size_t)nt, -1);
    std::vector<int> bad_code((size_t)nt, 0);
    const int64_t per = (n + nt - 1) / nt;
    auto run = [&](int t, int64_t hi_cap) {               // hi_cap < 0: the chunk's own end
        const int64_t lo = std::min(n, t * per), hi = hi_cap < 0 ? std::min(n, lo + per) : hi_cap;
        for (int64_t i = lo; i < hi; i++) {
            double nm = 0.0;
            const int code = normalize_row(elem_type, dim, (const uint8_t *)in + (size_t)i * rb, (uint8_t *)out + (size_t)i * rb, &nm);
            if (code) { first_bad[(size_t)t] = i; bad_code[(size_t)t] = code; return; }
            if (norms) norms[i] = nm;
        }
    };
    if (nt == 1) run(0, -1);
    else {
        std::vector<std::thread> ts;
        int started = 0;
        try {                                             // a thread that cannot be created leaves the rest to this one
            ts.reserve((size_t)nt);
            for (; started < nt; started++) ts.emplace_back(run, started, (int64_t)-1);
        } catch (...) {
        }
        if (started < nt) {                               // chunks started..nt-1 as one range, reported under slot `started`
            run(started, n);
        }
        for (auto &t : ts) t.join();
    }
    for (int t = 0; t < nt; t++)
        if (first_bad[(size_t)t] >= 0) return normalize_status(bad_code[(size_t)t]);
    return PGVRX_OK;
}

int pgvrx_index_create(const pgvrx_index_desc *d, pgvrx_index **out) {
    if (!d || !out) return fail(nullptr, PGVRX_ERR_INVALID, "null argument");
    *out = nullptr;
    // messages follow tests/pg_regress/expected/hnsw_vector.out:150-159
    // (the ERROR line; PostgreSQL's DETAIL line carries the valid range)
    if (d->m < PGVRX_MIN_M || d->m > PGVRX_MAX_M) {
        char b[96];
        snprintf(b, sizeof b, "value %d out of bounds for option \"m\"", d->m);
        return fail(nullptr, PGVRX_ERR_INVALID, b);
    }
    if (d->ef_construction < PGVRX_MIN_EF_CONSTRUCTION || d->ef_construction > PGVRX_MAX_EF_CONSTRUCTION) {
        char b[96];
        snprintf(b, sizeof b, "value %d out of bounds for option \"ef_construction\"", d->ef_construction);
        return fail(nullptr, PGVRX_ERR_INVALID, b);
    }
    if (d->ef_construction < 2 * d->m)
        return fail(nullptr, PGVRX_ERR_INVALID, "ef_construction must be greater than or equal to 2 * m");
    if (!metric_ok(d->elem_type, d->metric)) return fail(nullptr, PGVRX_ERR_INVALID, "operator class does not match the indexed type");
    if (d->dim < 1) return fail(nullptr, PGVRX_ERR_INVALID, "column does not have dimensions");
    if (d->dim > max_dim(d->elem_type)) {
        char b[128];
        snprintf(b, sizeof b, "column cannot have more than %d dimensions for hnsw index", max_dim(d->elem_type));
        return fail(nullptr, PGVRX_ERR_INVALID, b);
    }
    if (d->capacity < 0 || d->capacity > (int64_t)0x7ffffff0) return fail(nullptr, PGVRX_ERR_INVALID, "capacity out of range");
    int ndev = pgvrx_device_count();
    if (ndev <= 0) return fail(nullptr, PGVRX_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (d->device < 0 || d->device >= ndev) return fail(nullptr, PGVRX_ERR_INVALID, "bad device ordinal");
    pgvrx_index *ix = new (std::nothrow) pgvrx_index();
    if (!ix) return fail(nullptr, PGVRX_ERR_NOMEM, "out of host memory");
    ix->desc = *d;
    ix->row_bytes = pgvrx_row_bytes(d->elem_type, d->dim);
    ix->row_stride = round16u(ix->row_bytes);
    cudaError_t e = cudaSetDevice(d->device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, d->device);
    if (e == cudaSuccess && prop.major < 10) {
        delete ix;
        return fail(nullptr, PGVRX_ERR_UNSUPPORTED, "device is not sm_100-class (kernels are built for sm_100a only)");
    }
    if (e == cudaSuccess) ix->sm_count = prop.multiProcessorCount;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ix->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ix->ev1);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&ix->ev_ph[i]);
    if (e != cudaSuccess) {
        std::string msg = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e);
        pgvrx_index_destroy(ix);
        cudaGetLastError();
        return fail(nullptr, PGVRX_ERR_CUDA, msg);
    }
    *out = ix;
    return PGVRX_OK;
}

void pgvrx_index_destroy(pgvrx_index *ix) {
    if (!ix) return;
    cudaSetDevice(ix->desc.device);
    DevBuf *bufs[] = {&ix->vectors, &ix->adj0, &ix->adj_up, &ix->upper_base, &ix->level, &ix->flags, &ix->q, &ix->ids, &ix->dists,
                      &ix->counts,
                      &ix->part_ids, &ix->part_dists, &ix->cand, &ix->tc_xhi, &ix->tc_xlo, &ix->tc_xnorm, &ix->tc_xmax,
                      &ix->tc_qhi, &ix->tc_qlo, &ix->tc_qnorm, &ix->tc_cand, &ix->tc_bound, &ix->tc_uncert, &ix->tc_cnt,
                      &ix->tc_fq, &ix->tc_fids, &ix->tc_fdists, &ix->tc_dbg, &ix->tc_hint, &ix->tc_appcnt};
    for (DevBuf *b : bufs) b->release();
    for (auto &sc : ix->scan_scratch) {
        DevBuf *sb[] = {&sc.status, &sc.ctr, &sc.vis, &sc.vis_retry, &sc.qmap, &sc.work, &sc.phase};
        for (DevBuf *b : sb) b->release();
    }
    if (ix->build) pgvrx_host::destroy_build_state(ix->build);
    if (ix->ev0) cudaEventDestroy(ix->ev0);
    if (ix->ev1) cudaEventDestroy(ix->ev1);
    for (int i = 0; i < 4; i++) if (ix->ev_ph[i]) cudaEventDestroy(ix->ev_ph[i]);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

int64_t pgvrx_index_size(const pgvrx_index *ix) { return ix ? ix->n : 0; }

int pgvrx_index_load_graph(pgvrx_index *ix, const void *rows, int64_t n, const uint8_t *level, const int32_t *adj0,
                           const int32_t *upper_base, const int32_t *adj_up, int64_t n_upper_rows, int64_t entry) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (n < 0 || n > ix->desc.capacity) return fail(ix, PGVRX_ERR_INVALID, "n exceeds the handle's capacity");
    if (n > 0 && (!rows || !level || !adj0 || !upper_base)) return fail(ix, PGVRX_ERR_INVALID, "null graph array");
    if (n_upper_rows < 0 || (n_upper_rows > 0 && !adj_up)) return fail(ix, PGVRX_ERR_INVALID, "bad upper-layer rows");
    if (n == 0) entry = -1;
    if (n > 0 && (entry < 0 || entry >= n)) return fail(ix, PGVRX_ERR_INVALID, "entry point out of range");
    const int m = ix->desc.m;
    // validate ids once on the host: the kernels trust them
    for (int64_t i = 0; i < n; i++) {
        int lv = level[i];
        if ((lv > 0) != (upper_base[i] >= 0) || (lv > 0 && upper_base[i] + lv > n_upper_rows))
            return fail(ix, PGVRX_ERR_INVALID, "upper_base inconsistent with level");
        if (lv > pgvrx_max_level(m)) return fail(ix, PGVRX_ERR_INVALID, "level above the maximum for m");
    }
    for (int64_t i = 0; i < n * 2 * m; i++)
        if (adj0[i] < -1 || adj0[i] >= n) return fail(ix, PGVRX_ERR_INVALID, "layer-0 neighbour id out of range");
    for (int64_t i = 0; i < n_upper_rows * m; i++)
        if (adj_up[i] < -1 || adj_up[i] >= n) return fail(ix, PGVRX_ERR_INVALID, "upper-layer neighbour id out of range");

    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    ix->generation++;
    ix->loaded = false;
    ix->rows_loaded = false;
    ix->tc_ready = false;
    ix->has_flags = false;
    if (ix->build) { pgvrx_host::destroy_build_state(ix->build); ix->build = nullptr; }   // the host mirror described the old graph
    CU_TRY(ix, ix->vectors.ensure(std::max<size_t>(16, (size_t)n * ix->row_stride)));
    CU_TRY(ix, ix->adj0.ensure(std::max<size_t>(16, (size_t)n * 2 * m * 4)));
    CU_TRY(ix, ix->adj_up.ensure(std::max<size_t>(16, (size_t)n_upper_rows * m * 4)));
    CU_TRY(ix, ix->upper_base.ensure(std::max<size_t>(16, (size_t)n * 4)));
    CU_TRY(ix, ix->level.ensure(std::max<size_t>(16, (size_t)n)));
    if (n > 0) {
        if (ix->row_stride == ix->row_bytes) {
            CU_TRY(ix, cudaMemcpyAsync(ix->vectors.p, rows, (size_t)n * ix->row_bytes, cudaMemcpyHostToDevice, ix->stream));
        } else {
            CU_TRY(ix, cudaMemsetAsync(ix->vectors.p, 0, (size_t)n * ix->row_stride, ix->stream));
            CU_TRY(ix, cudaMemcpy2DAsync(ix->vectors.p, ix->row_stride, rows, ix->row_bytes, ix->row_bytes, (size_t)n,
                                         cudaMemcpyHostToDevice, ix->stream));
        }
        CU_TRY(ix, cudaMemcpyAsync(ix->adj0.p, adj0, (size_t)n * 2 * m * 4, cudaMemcpyHostToDevice, ix->stream));
        if (n_upper_rows > 0)
            CU_TRY(ix, cudaMemcpyAsync(ix->adj_up.p, adj_up, (size_t)n_upper_rows * m * 4, cudaMemcpyHostToDevice, ix->stream));
        CU_TRY(ix, cudaMemcpyAsync(ix->upper_base.p, upper_base, (size_t)n * 4, cudaMemcpyHostToDevice, ix->stream));
        CU_TRY(ix, cudaMemcpyAsync(ix->level.p, level, (size_t)n, cudaMemcpyHostToDevice, ix->stream));
        CU_TRY(ix, cudaStreamSynchronize(ix->stream));
    }
    ix->n = n;
    ix->n_upper = n_upper_rows;
    ix->entry = entry;
    ix->entry_level = n > 0 ? level[entry] : -1;
    ix->loaded = true;
    ix->rows_loaded = true;
    return PGVRX_OK;
}

int pgvrx_index_set_element_flags(pgvrx_index *ix, const uint8_t *flags, int64_t n) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (!ix->loaded) return fail(ix, PGVRX_ERR_STATE, "no graph loaded");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    // scans in flight on caller streams read the old flags: drain them before the table changes
    CU_TRY(ix, cudaDeviceSynchronize());
    ix->generation++;
    if (!flags) { ix->has_flags = false; return PGVRX_OK; }
    if (n != ix->n) return fail(ix, PGVRX_ERR_INVALID, "flags must cover every element of the index");
    bool any = false;
    for (int64_t i = 0; i < n; i++) {
        if (flags[i] & ~(PGVRX_ELEM_DELETED | PGVRX_ELEM_NO_HEAPTIDS)) return fail(ix, PGVRX_ERR_INVALID, "unknown element flag");
        any = any || flags[i] != 0;
    }
    ix->has_flags = false;
    if (!any) return PGVRX_OK;                       // nothing flagged: keep the kernels on the flag-free path
    CU_TRY(ix, ix->flags.ensure(std::max<size_t>(16, (size_t)n)));
    CU_TRY(ix, cudaMemcpy(ix->flags.p, flags, (size_t)n, cudaMemcpyHostToDevice));
    ix->has_flags = true;
    return PGVRX_OK;
}

int pgvrx_search_batch_device(pgvrx_index *ix, const void *d_queries, int64_t nq, int32_t ef_search, int32_t k,
                              int32_t *d_ids, float *d_dists, int32_t *d_counts, pgvrx_search_stats *stats,
                              void *cuda_stream) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (nq > 0 && (!d_queries || !d_ids || !d_dists || !d_counts)) return fail(ix, PGVRX_ERR_INVALID, "null device buffer");
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
    return search_core(ix, d_queries, (uint32_t)ix->row_bytes, nq, ef_search, k, d_ids, d_dists, d_counts, stats, s, nullptr, nullptr);
}

// device-visible alias of a host buffer when it is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory):
// the kernels then read queries and write results through it directly (each query row is read once, each result
// row written once), so the call has no separate copy phases
static void *mapped_alias(const void *host_ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return at.devicePointer;
}

int pgvrx_search_batch(pgvrx_index *ix, const void *queries, int64_t nq, int32_t ef_search, int32_t k, int32_t *out_ids,
                       float *out_dists, int32_t *out_counts, pgvrx_search_stats *stats) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (nq > 0 && (!queries || !out_ids || !out_dists || !out_counts)) return fail(ix, PGVRX_ERR_INVALID, "null buffer");
    if (nq < 0 || k < 1) return fail(ix, PGVRX_ERR_INVALID, "bad k or nq");
    if (nq == 0) { if (stats) memset(stats, 0, sizeof *stats); return PGVRX_OK; }
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    // page-locked caller buffers are used in place; pageable ones are staged through device copies
    const bool zc_ok = getenv("PGVRX_NO_ZEROCOPY") == nullptr && (ix->row_bytes % 16) == 0;
    void *zq = zc_ok ? mapped_alias(queries) : nullptr;
    void *zi = zc_ok ? mapped_alias(out_ids) : nullptr, *zd = zc_ok ? mapped_alias(out_dists) : nullptr,
         *zcnt = zc_ok ? mapped_alias(out_counts) : nullptr;
    const bool zc_out = zi && zd && zcnt;
    if (!zq) {
        CU_TRY(ix, ix->q.ensure((size_t)nq * ix->row_bytes));
        CU_TRY(ix, cudaMemcpyAsync(ix->q.p, queries, (size_t)nq * ix->row_bytes, cudaMemcpyHostToDevice, ix->stream));
    }
    if (!zc_out) {
        CU_TRY(ix, ix->ids.ensure((size_t)nq * k * 4));
        CU_TRY(ix, ix->dists.ensure((size_t)nq * k * 4));
        CU_TRY(ix, ix->counts.ensure((size_t)nq * 4));
    }
    int64_t n_failed = 0, first_failed = -1;
    int rc = search_core(ix, zq ? zq : ix->q.p, (uint32_t)ix->row_bytes, nq, ef_search, k, zc_out ? (int32_t *)zi : (int32_t *)ix->ids.p,
                         zc_out ? (float *)zd : (float *)ix->dists.p, zc_out ? (int32_t *)zcnt : (int32_t *)ix->counts.p, stats,
                         ix->stream, &n_failed, &first_failed);
    if (rc) return rc;
    if (!zc_out) {
        CU_TRY(ix, cudaMemcpyAsync(out_ids, ix->ids.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
        CU_TRY(ix, cudaMemcpyAsync(out_dists, ix->dists.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
        CU_TRY(ix, cudaMemcpyAsync(out_counts, ix->counts.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ix->stream));
        CU_TRY(ix, cudaStreamSynchronize(ix->stream));
    }
    if (n_failed > 0) {
        // every other query's rows are in the caller's arrays; the failed ones carry out_counts = -1
        char b[200];
        snprintf(b, sizeof b, "%lld of %lld queries exhausted the search scratch even in the retry pass (first: query %lld); "
                              "their out_counts are -1, the other results are complete",
                 (long long)n_failed, (long long)nq, (long long)first_failed);
        return fail(ix, PGVRX_ERR_UNSUPPORTED, b);
    }
    return PGVRX_OK;
}

}  // extern "C"
