// flat_tc_host.cu -- orchestration of the tensor-core flat path (K4b): operand preparation, TMA tensor maps,
// the 3xTF32 tcgen05 candidate kernel, the exact re-rank, certification against the error bound, and the
// exact re-run of the (rare) queries that could not be certified.  See flat_tc.cu for the kernel.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "index_internal.h"

using namespace pgvrx;

int pgvrx_flat_exact_core(pgvrx_index *ix, const void *d_queries, int64_t nq, int k, int64_t row_begin, int64_t row_end,
                          int64_t id_base, int32_t *d_ids, float *d_dists, pgvrx_flat_stats *stats, cudaStream_t stream);

namespace {

typedef CUresult (*encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_t get_encode() {
    static encode_tiled_t fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) {
        cudaGetLastError();
        return nullptr;
    }
    fn = (encode_tiled_t)p;
    return fn;
}

// rows x d_pad fp32, K-major; box = 128 rows x 32 floats (one 128-byte swizzle row per matrix row)
int make_map(pgvrx_index *ix, void *out128, void *base, int64_t rows, int d_pad) {
    encode_tiled_t enc = get_encode();
    if (!enc) return pgvrx_fail(ix, PGVRX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)d_pad * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_M};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(reinterpret_cast<CUtensorMap *>(out128), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[96];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return pgvrx_fail(ix, PGVRX_ERR_CUDA, b);
    }
    return PGVRX_OK;
}

// |approx - reference| <= eps_rel * (|q|^2 + max|x|^2), DESIGN.md section 4 (K4b): splitting residuals 3 * 2^-22,
// fp32 accumulation of up to 3 d products in the tensor core and d in the reference chain, doubled for safety
float tc_eps_rel(int dim) { return (float)((dim + 2) * std::ldexp(1.0, -21)); }

int prepare_table(pgvrx_index *ix, cudaStream_t stream) {
    if (ix->tc_ready) return PGVRX_OK;
    const int dim = ix->desc.dim;
    const int d_pad = (dim + TC_KB - 1) / TC_KB * TC_KB;
    const int64_t n_pad = std::max<int64_t>(2 * TC_N, (ix->n + 2 * TC_N - 1) / (2 * TC_N) * (2 * TC_N));   // whole 256-row tiles (2-CTA kernel)
    CU_TRY(ix, ix->tc_xhi.ensure((size_t)n_pad * d_pad * 4));
    CU_TRY(ix, ix->tc_xlo.ensure((size_t)n_pad * d_pad * 4));
    CU_TRY(ix, ix->tc_xnorm.ensure((size_t)n_pad * 4));
    CU_TRY(ix, ix->tc_xmax.ensure(64));
    CU_TRY(ix, cudaMemsetAsync(ix->tc_xmax.p, 0, 64, stream));
    CU_TRY(ix, launch_tc_split((const uint8_t *)ix->vectors.p, ix->row_stride, ix->n, n_pad, dim, d_pad, (float *)ix->tc_xhi.p,
                               (float *)ix->tc_xlo.p, (float *)ix->tc_xnorm.p, (uint32_t *)ix->tc_xmax.p, nullptr, stream));
    int rc = make_map(ix, ix->tc_map_xhi, ix->tc_xhi.p, n_pad, d_pad);
    if (rc) return rc;
    rc = make_map(ix, ix->tc_map_xlo, ix->tc_xlo.p, n_pad, d_pad);
    if (rc) return rc;
    ix->tc_n_pad = n_pad;
    ix->tc_d_pad = d_pad;
    ix->tc_ready = true;
    return PGVRX_OK;
}

struct TcPlan { int n_mtiles, n_ntiles, n_splits, tiles_per_split, grid; };

// tile_m x tile_n = 128 x 128 per CTA, or 256 x 256 per CTA pair (`workers` = CTAs or pairs)
TcPlan plan_tc(int workers, int64_t nq_pad, int64_t n_pad, int tile_m, int tile_n, int split_penalty = 10) {
    const int sm_count = workers;
    TcPlan pl;
    pl.n_mtiles = (int)(nq_pad / tile_m);
    pl.n_ntiles = (int)(n_pad / tile_n);
    // units = query tiles x base splits are dealt round-robin to one CTA per SM: pick the split count with the
    // least total time (waves x tiles per unit); more splits only add candidates to re-rank
    long best_cost = -1;
    pl.n_splits = 1;
    pl.tiles_per_split = pl.n_ntiles;
    for (int s = 1; s <= std::min(pl.n_ntiles, 64); s++) {
        const int tps = (pl.n_ntiles + s - 1) / s;
        const int s_eff = (pl.n_ntiles + tps - 1) / tps;
        const long units = (long)pl.n_mtiles * s_eff;
        const long waves = (units + sm_count - 1) / sm_count;
        const long cost = waves * tps + (long)split_penalty * s_eff;   // each split adds up to KP candidates per query to the re-rank
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; pl.n_splits = s_eff; pl.tiles_per_split = tps; }
    }
    pl.grid = (int)std::min<long>((long)pl.n_mtiles * pl.n_splits, sm_count);
    return pl;
}

}  // namespace

bool pgvrx_flat_tc_eligible(const pgvrx_index *ix, int64_t nq, int k) {
    if (ix->desc.elem_type != PGVRX_ELEM_F32) return false;
    if (ix->desc.metric != PGVRX_METRIC_L2SQ && ix->desc.metric != PGVRX_METRIC_NEG_IP) return false;
    if (k > TC_MAX_K || ix->n < 1 || nq < 1) return false;
    return true;
}

// candidates -> exact top-k -> certification -> exact re-run of what could not be certified
int pgvrx_flat_tc_core(pgvrx_index *ix, const void *d_queries, int64_t nq, int k, int64_t id_base, int32_t *d_ids, float *d_dists,
                       pgvrx_flat_stats *stats, float *d_dbg_scores, cudaStream_t stream) {
    if (id_base < 0 || id_base + ix->n > (int64_t)0x7fffffff) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "ids exceed int32");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    if (stats) memset(stats, 0, sizeof *stats);
    CU_TRY(ix, cudaEventRecord(ix->ev0, stream));
    int rc = prepare_table(ix, stream);
    if (rc) return rc;
    const int d_pad = ix->tc_d_pad;
    // CTA pairs (tcgen05 cta_group::2) when the query tile can stay resident; PGVRX_TC_1CTA pins the single-CTA kernel
    const bool two_cta = d_pad <= 4 * TC_KB && !getenv("PGVRX_TC_1CTA") && !getenv("PGVRX_TC_STREAM_Q") && (ix->sm_count % 2) == 0;
    const int tile_m = two_cta ? 2 * TC_M : TC_M, tile_n = two_cta ? 2 * TC_N : TC_N;
    const int64_t nq_pad = (nq + tile_m - 1) / tile_m * tile_m;
    CU_TRY(ix, ix->tc_qhi.ensure((size_t)nq_pad * d_pad * 4));
    CU_TRY(ix, ix->tc_qlo.ensure((size_t)nq_pad * d_pad * 4));
    CU_TRY(ix, ix->tc_qnorm.ensure((size_t)nq_pad * 4));
    CU_TRY(ix, ix->tc_hint.ensure((size_t)nq_pad * 4));
    CU_TRY(ix, launch_tc_split((const uint8_t *)d_queries, (uint32_t)ix->row_bytes, nq, nq_pad, ix->desc.dim, d_pad,
                               (float *)ix->tc_qhi.p, (float *)ix->tc_qlo.p, (float *)ix->tc_qnorm.p, nullptr,
                               (uint32_t *)ix->tc_hint.p, stream));
    if (stats) CU_TRY(ix, cudaEventRecord(ix->ev_ph[0], stream));
    alignas(64) unsigned char map_qhi[128], map_qlo[128];
    rc = make_map(ix, map_qhi, ix->tc_qhi.p, nq_pad, d_pad);
    if (rc) return rc;
    rc = make_map(ix, map_qlo, ix->tc_qlo.p, nq_pad, d_pad);
    if (rc) return rc;

    TcPlan pl = plan_tc(two_cta ? ix->sm_count / 2 : ix->sm_count, nq_pad, ix->tc_n_pad, tile_m, tile_n);
    // short units get a threshold from the sample pre-pass (below): what a query inserts is then the rows under that
    // threshold, however the table is split, so splits cost little and the plan may balance the waves more finely
    // (the pre-pass runs one unit per query tile -- 40 of the 74 CTA pairs for 10k queries -- over 1/24 of the table: worth
    // it up to about a million 128-d rows per device (1M: 11.45 -> 10.27 ms), a loss beyond (4M x 96: 38.1 -> 41.6 ms))
    int max_tps = 512;
    if (const char *e = getenv("PGVRX_TC_PREPASS_MAX_TILES")) max_tps = atoi(e);
    const bool want_prepass = d_ids && pl.tiles_per_split <= max_tps && pl.n_ntiles >= 64 && !getenv("PGVRX_TC_NO_PREPASS");
    if (want_prepass) pl = plan_tc(two_cta ? ix->sm_count / 2 : ix->sm_count, nq_pad, ix->tc_n_pad, tile_m, tile_n, 3);
    const int nc = pl.n_splits * TC_KP;
    CU_TRY(ix, ix->tc_cand.ensure((size_t)nq * nc * 4));
    CU_TRY(ix, ix->tc_bound.ensure((size_t)pl.n_splits * nq * 4));
    CU_TRY(ix, ix->tc_uncert.ensure((size_t)nq * 4));
    CU_TRY(ix, ix->tc_cnt.ensure(64));
    CU_TRY(ix, cudaMemsetAsync(ix->tc_cnt.p, 0, 64, stream));

    TcParams p;
    p.nq = (int)nq;
    p.n_mtiles = pl.n_mtiles;
    p.n_rows = ix->n;
    p.n_ntiles = pl.n_ntiles;
    p.n_splits = pl.n_splits;
    p.ntiles_per_split = pl.tiles_per_split;
    p.d_pad = d_pad;
    p.kp = TC_KP;
    p.metric = ix->desc.metric;
    p.a_resident = (d_pad <= 4 * TC_KB && !getenv("PGVRX_TC_STREAM_Q")) ? 1 : 0;
    p.id_base = id_base;
    p.qnorm = (const float *)ix->tc_qnorm.p;
    p.xnorm = (const float *)ix->tc_xnorm.p;
    p.cand_ids = (int32_t *)ix->tc_cand.p;
    p.cand_bound = (float *)ix->tc_bound.p;
    p.dbg_scores = d_dbg_scores;
    p.hint = (uint32_t *)ix->tc_hint.p;
    p.tile_step = 1;
    p.kp_eff = TC_KP;
    p.local_k = 0;
    p.app_ids = nullptr;
    p.app_cnt = nullptr;
    p.app_cap = 0;
    // Hint pre-pass.  A unit (query tile x base split) starts with empty candidate sets, and until a set's maximum has
    // come down nearly every element takes the serial insertion path: KP (1 + ln(rows / KP)) insertions per query and
    // unit.  Units of later waves inherit a threshold from earlier ones through p.hint; the first wave does not, and
    // with short units (a shard of a row-sharded table: 125k rows = 98 tiles per unit) that warm-up is as long as the
    // tensor work itself (measured: 3.3 ms for 1.65 ms of MMAs).  So every query first sees a strided SAMPLE of the
    // table -- every step-th tile, one unit per query tile, sets of kp_eff keys -- which leaves in p.hint the kp_eff-th
    // smallest approximate distance of the sample: about kp_eff * step rows of the table lie below it, several times k.
    // The main pass then starts from that threshold everywhere.  The certificate takes the final hint into account.
    const bool prepass = want_prepass;
    int app_cap = 0;
    if (prepass) {
        TcParams pa = p;
        // a 1/24 sample and the (k / 2 + 3)-th smallest of it: ~ 24 (k / 2 + 3) rows of the table below the threshold
        // (190 for k = 10), an order of magnitude more than the k the certificate needs, a fraction of what unguided
        // candidate sets would insert
        const int want_tiles = std::max(4, pl.n_ntiles / 24);
        pa.tile_step = std::max(1, pl.n_ntiles / want_tiles);
        pa.kp_eff = std::max(8, std::min((int)TC_KP, k / 2 + 3));
        // lane-local top lists of 8 (k <= 10) or 16 keys; PGVRX_TC_PREPASS_SETS=1 keeps the set-maintaining form (kp_eff keys)
        if (!getenv("PGVRX_TC_PREPASS_SETS")) { pa.local_k = k <= 10 ? 8 : 16; pa.kp_eff = pa.local_k; }
        pa.n_splits = 1;
        pa.ntiles_per_split = pl.n_ntiles;
        pa.cand_ids = nullptr;
        pa.cand_bound = nullptr;
        pa.dbg_scores = nullptr;
        const int workers = two_cta ? ix->sm_count / 2 : ix->sm_count;
        const int ga = std::min(pl.n_mtiles, workers);
        if (two_cta) CU_TRY(ix, launch_flat_tc2(map_qhi, map_qlo, ix->tc_map_xhi, ix->tc_map_xlo, pa, 2 * ga, stream));
        else CU_TRY(ix, launch_flat_tc(map_qhi, map_qlo, ix->tc_map_xhi, ix->tc_map_xlo, pa, ga, stream));
        // Main pass in append mode: with a threshold known up front nothing has to be maintained -- a lane appends its own
        // query's rows under the threshold (expected kp_eff * step of them per query, Gamma distributed; the list holds
        // five times that) and the serial set insertions disappear from the epilogue.  PGVRX_TC_NO_APPEND=1 keeps the sets.
        if (!getenv("PGVRX_TC_NO_APPEND")) {
            // one segment per (query, split): expected kp_eff * step / n_splits rows each, room for five times that + 32
            app_cap = (int)pgvrx_pow2_at_least((uint64_t)(5 * pa.kp_eff * pa.tile_step / pl.n_splits + 32));
            app_cap = std::max(64, std::min(app_cap, 4096));
            CU_TRY(ix, ix->tc_cand.ensure((size_t)nq * pl.n_splits * app_cap * 4));
            CU_TRY(ix, ix->tc_appcnt.ensure((size_t)nq * pl.n_splits * 4));
            CU_TRY(ix, cudaMemsetAsync(ix->tc_cand.p, 0xff, (size_t)nq * pl.n_splits * app_cap * 4, stream));
            CU_TRY(ix, cudaMemsetAsync(ix->tc_appcnt.p, 0, (size_t)nq * pl.n_splits * 4, stream));
            p.app_ids = (int32_t *)ix->tc_cand.p;
            p.app_cnt = (int *)ix->tc_appcnt.p;
            p.app_cap = app_cap;
            p.cand_ids = nullptr;
            p.cand_bound = nullptr;
        }
    }
    if (two_cta) CU_TRY(ix, launch_flat_tc2(map_qhi, map_qlo, ix->tc_map_xhi, ix->tc_map_xlo, p, 2 * pl.grid, stream));
    else CU_TRY(ix, launch_flat_tc(map_qhi, map_qlo, ix->tc_map_xhi, ix->tc_map_xlo, p, pl.grid, stream));
    int launches = prepass ? 3 : 2;
    uint64_t n_uncert = 0;
    if (stats) CU_TRY(ix, cudaEventRecord(ix->ev_ph[1], stream));
    if (d_ids) {
        rc = pgvrx_rerank_device(ix, d_queries, nq, (const int32_t *)ix->tc_cand.p, app_cap ? app_cap * pl.n_splits : nc, k, id_base, d_ids, d_dists, stream);
        if (rc) return rc;
        if (stats) CU_TRY(ix, cudaEventRecord(ix->ev_ph[2], stream));
        CU_TRY(ix, launch_tc_certify(d_dists, d_ids, k, (const float *)ix->tc_bound.p, app_cap ? 0 : pl.n_splits, (int)nq,
                                     (const float *)ix->tc_qnorm.p, (const uint32_t *)ix->tc_xmax.p, tc_eps_rel(ix->desc.dim),
                                     ix->desc.metric, (const uint32_t *)ix->tc_hint.p, app_cap ? (const int *)ix->tc_appcnt.p : nullptr, app_cap, pl.n_splits,
                                     (int32_t *)ix->tc_uncert.p, (int *)ix->tc_cnt.p, stream));
        launches += 2;
        int nu = 0;
        CU_TRY(ix, cudaMemcpyAsync(&nu, ix->tc_cnt.p, 4, cudaMemcpyDeviceToHost, stream));
        CU_TRY(ix, cudaStreamSynchronize(stream));
        n_uncert = (uint64_t)nu;
        if (nu > 0) {
            // the bound could not separate these queries' k-th neighbour from the unkept rows: exact scan for them
            CU_TRY(ix, ix->tc_fq.ensure((size_t)nu * ix->row_bytes));
            CU_TRY(ix, ix->tc_fids.ensure((size_t)nu * k * 4));
            CU_TRY(ix, ix->tc_fdists.ensure((size_t)nu * k * 4));
            CU_TRY(ix, launch_tc_gather_queries((const uint8_t *)d_queries, (uint32_t)ix->row_bytes, (const int32_t *)ix->tc_uncert.p, nu,
                                                (uint8_t *)ix->tc_fq.p, stream));
            rc = pgvrx_flat_exact_core(ix, ix->tc_fq.p, nu, k, 0, ix->n, id_base, (int32_t *)ix->tc_fids.p, (float *)ix->tc_fdists.p,
                                       nullptr, stream);
            if (rc) return rc;
            CU_TRY(ix, launch_tc_scatter_results((const int32_t *)ix->tc_fids.p, (const float *)ix->tc_fdists.p, k,
                                                 (const int32_t *)ix->tc_uncert.p, nu, d_ids, d_dists, stream));
            launches += 4;
        }
    }
    CU_TRY(ix, cudaEventRecord(ix->ev1, stream));
    CU_TRY(ix, cudaStreamSynchronize(stream));
    if (stats) {
        CU_TRY(ix, cudaEventElapsedTime(&stats->kernel_ms, ix->ev0, ix->ev1));
        stats->n_pairs = (uint64_t)nq * (uint64_t)ix->n;
        stats->n_candidates = (uint64_t)nq * (uint64_t)nc;
        if (app_cap) {                                   // append mode: what the lists really hold
            std::vector<int> cnt((size_t)nq * pl.n_splits);
            CU_TRY(ix, cudaMemcpy(cnt.data(), ix->tc_appcnt.p, cnt.size() * 4, cudaMemcpyDeviceToHost));
            uint64_t tot = 0;
            for (int c : cnt) tot += (uint64_t)std::min(c, app_cap);
            stats->n_candidates = tot;
        }
        stats->n_uncertified = n_uncert;
        stats->n_launches = (uint32_t)launches;
        stats->n_splits = (uint32_t)pl.n_splits;
        stats->path = two_cta ? 2 : 1;
        CU_TRY(ix, cudaEventElapsedTime(&stats->phase_ms[0], ix->ev0, ix->ev_ph[0]));
        CU_TRY(ix, cudaEventElapsedTime(&stats->phase_ms[1], ix->ev_ph[0], ix->ev_ph[1]));
        if (d_ids) {
            CU_TRY(ix, cudaEventElapsedTime(&stats->phase_ms[2], ix->ev_ph[1], ix->ev_ph[2]));
            CU_TRY(ix, cudaEventElapsedTime(&stats->phase_ms[3], ix->ev_ph[2], ix->ev1));
        }
    }
    return PGVRX_OK;
}

extern "C" {

// Diagnostic: the approximate distances the tensor-core path ranks by (nq x n, HOST buffers) and the bound
// eps_rel * (|q|^2 + max |x|^2) it certifies against.  Lets a test measure the real 3xTF32 error against the bound.
int pgvrx_flat_tc_scores(pgvrx_index *ix, const void *queries, int64_t nq, float *out_scores, float *out_eps) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (!ix->rows_loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no rows loaded");
    if (!pgvrx_flat_tc_eligible(ix, nq, 1) || !queries || !out_scores) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "tensor-core path needs fp32 L2 / inner product");
    if ((uint64_t)nq * (uint64_t)ix->n > ((uint64_t)1 << 28)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "score dump limited to 2^28 pairs");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    CU_TRY(ix, ix->q.ensure((size_t)nq * ix->row_bytes));
    CU_TRY(ix, ix->tc_dbg.ensure((size_t)nq * ix->n * 4));
    CU_TRY(ix, cudaMemcpyAsync(ix->q.p, queries, (size_t)nq * ix->row_bytes, cudaMemcpyHostToDevice, ix->stream));
    int rc = pgvrx_flat_tc_core(ix, ix->q.p, nq, 1, 0, nullptr, nullptr, nullptr, (float *)ix->tc_dbg.p, ix->stream);
    if (rc) return rc;
    CU_TRY(ix, cudaMemcpy(out_scores, ix->tc_dbg.p, (size_t)nq * ix->n * 4, cudaMemcpyDeviceToHost));
    if (out_eps) {
        std::vector<float> qn((size_t)nq);
        uint32_t xm = 0;
        CU_TRY(ix, cudaMemcpy(qn.data(), ix->tc_qnorm.p, (size_t)nq * 4, cudaMemcpyDeviceToHost));
        CU_TRY(ix, cudaMemcpy(&xm, ix->tc_xmax.p, 4, cudaMemcpyDeviceToHost));
        float xmf;
        memcpy(&xmf, &xm, 4);
        for (int64_t i = 0; i < nq; i++) out_eps[i] = tc_eps_rel(ix->desc.dim) * (qn[(size_t)i] + xmf);
    }
    return PGVRX_OK;
}

}  // extern "C"
