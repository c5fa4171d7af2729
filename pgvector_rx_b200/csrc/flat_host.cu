// flat_host.cu -- C-ABI entry points of the exact paths: flat kNN over the vector table (K4), exact
// re-rank of candidate lists, and the k-way top-k merge (K6) used after an all-gather of per-shard results.
// No distance is computed on the host here; without a device every call fails (no CPU path).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "index_internal.h"

using namespace pgvrx;

namespace {

struct FlatPlan {
    FlatGeom geom;
    size_t smem;
    int n_qtiles, n_splits;
    int64_t rows_per_split;
};

int flat_plan(pgvrx_index *ix, int64_t nq, int k, int64_t n_rows, FlatPlan *out) {
    FlatPlan pl;
    pl.smem = flat_smem_bytes(ix->desc.elem_type, ix->row_stride, k, &pl.geom);
    if (pl.smem > 227 * 1024) return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "k too large for the flat kernel's shared memory");
    int occ = 0;
    CU_TRY(ix, flat_kernel_occupancy(ix->desc.elem_type, ix->desc.metric, pl.smem, &occ));
    if (occ < 1) return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "flat kernel does not fit on an SM");
    pl.n_qtiles = (int)((nq + FLAT_QT - 1) / FLAT_QT);
    const int64_t n_tiles = std::max<int64_t>(1, (n_rows + FLAT_THREADS - 1) / FLAT_THREADS);
    // at least four waves of CTAs so the last wave's tail stays small, never more splits than row tiles
    const int64_t slots = (int64_t)ix->sm_count * occ;
    int64_t want = (4 * slots + pl.n_qtiles - 1) / pl.n_qtiles;
    want = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, n_tiles), 4096));
    const int64_t tiles_per_split = (n_tiles + want - 1) / want;
    pl.rows_per_split = tiles_per_split * FLAT_THREADS;
    pl.n_splits = (int)((n_tiles + tiles_per_split - 1) / tiles_per_split);
    *out = pl;
    return PGVRX_OK;
}

}  // namespace

bool pgvrx_flat_tc_eligible(const pgvrx_index *ix, int64_t nq, int k);
int pgvrx_flat_tc_core(pgvrx_index *ix, const void *d_queries, int64_t nq, int k, int64_t id_base, int32_t *d_ids, float *d_dists,
                       pgvrx_flat_stats *stats, float *d_dbg_scores, cudaStream_t stream);

// K4a: sequential f32 chains on the CUDA cores (every opclass, any k up to 1024, any row range)
int pgvrx_flat_exact_core(pgvrx_index *ix, const void *d_queries, int64_t nq, int k, int64_t row_begin, int64_t row_end,
                          int64_t id_base, int32_t *d_ids, float *d_dists, pgvrx_flat_stats *stats, cudaStream_t stream) {
    if (!ix->rows_loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no rows loaded");
    if (k < 1 || k > FLAT_MAX_K) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "k must be between 1 and 1024");
    if (nq < 0 || nq > (int64_t)1 << 30) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad nq");
    if (row_begin < 0 || row_end > ix->n || row_begin > row_end) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "row range out of bounds");
    if (id_base < 0 || id_base + (row_end - row_begin) > (int64_t)0x7fffffff) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "ids exceed int32");
    if (stats) memset(stats, 0, sizeof *stats);
    if (nq == 0) return PGVRX_OK;
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    const int64_t n_rows = row_end - row_begin;
    FlatPlan pl;
    int rc = flat_plan(ix, nq, k, n_rows, &pl);
    if (rc) return rc;
    if (stats) CU_TRY(ix, cudaEventRecord(ix->ev0, stream));
    int launches = 0;
    if (n_rows == 0) {
        // ORDER BY over an empty table: no rows
        CU_TRY(ix, cudaMemsetAsync(d_ids, 0xff, (size_t)nq * k * 4, stream));
        std::vector<float> inf((size_t)nq * k, std::numeric_limits<float>::infinity());
        CU_TRY(ix, cudaMemcpyAsync(d_dists, inf.data(), inf.size() * 4, cudaMemcpyHostToDevice, stream));
        CU_TRY(ix, cudaStreamSynchronize(stream));
    } else {
        FlatParams p;
        p.vectors = (const uint8_t *)ix->vectors.p;
        p.row_stride = ix->row_stride;
        p.row_begin = row_begin;
        p.row_end = row_end;
        p.rows_per_split = pl.rows_per_split;
        p.id_base = id_base;
        p.queries = (const uint8_t *)d_queries;
        p.q_stride = (uint32_t)ix->row_bytes;
        p.q_bytes = (uint32_t)ix->row_bytes;
        p.nq = (int)nq;
        p.k = k;
        p.flags = ix->has_flags ? (const uint8_t *)ix->flags.p : nullptr;
        if (pl.n_splits == 1) {
            p.out_ids = d_ids;
            p.out_dists = d_dists;
        } else {
            CU_TRY(ix, ix->part_ids.ensure((size_t)pl.n_splits * nq * k * 4));
            CU_TRY(ix, ix->part_dists.ensure((size_t)pl.n_splits * nq * k * 4));
            p.out_ids = (int32_t *)ix->part_ids.p;
            p.out_dists = (float *)ix->part_dists.p;
        }
        CU_TRY(ix, launch_flat_partial(ix->desc.elem_type, ix->desc.metric, p, pl.geom, pl.n_qtiles, pl.n_splits, pl.smem, stream));
        launches++;
        if (pl.n_splits > 1) {
            CU_TRY(ix, launch_topk_merge(p.out_ids, p.out_dists, pl.n_splits, (int)nq, k, k, d_ids, d_dists, stream));
            launches++;
        }
    }
    if (stats) {
        CU_TRY(ix, cudaEventRecord(ix->ev1, stream));
        CU_TRY(ix, cudaStreamSynchronize(stream));
        CU_TRY(ix, cudaEventElapsedTime(&stats->kernel_ms, ix->ev0, ix->ev1));
        stats->n_pairs = (uint64_t)nq * (uint64_t)n_rows;
        stats->n_launches = (uint32_t)launches;
        stats->n_splits = (uint32_t)pl.n_splits;
        stats->path = 0;
    }
    return PGVRX_OK;
}

namespace {

// Which kernel serves a flat request.  The tensor cores take large fp32 L2 / inner-product batches over the
// whole table (north star item 3); everything else -- and whatever the error bound cannot certify -- runs on the
// exact CUDA-core kernel.  PGVRX_FLAT_PATH=exact|tc pins the choice (tests, benchmarks).
int flat_core(pgvrx_index *ix, const void *d_queries, int64_t nq, int k, int64_t row_begin, int64_t row_end, int64_t id_base,
              int32_t *d_ids, float *d_dists, pgvrx_flat_stats *stats, cudaStream_t stream) {
    if (!ix->rows_loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no rows loaded");
    // (an index that carries vacuum flags takes the CUDA-core kernel, which skips the flagged rows)
    bool tc = row_begin == 0 && row_end == ix->n && k >= 1 && !ix->has_flags && pgvrx_flat_tc_eligible(ix, nq, k);
    const char *pin = getenv("PGVRX_FLAT_PATH");
    if (pin && !strcmp(pin, "exact")) tc = false;
    else if (!(pin && !strcmp(pin, "tc"))) tc = tc && nq >= 512 && ix->n >= 16384;
    if (ix->has_flags) tc = false;
    if (tc) return pgvrx_flat_tc_core(ix, d_queries, nq, k, id_base, d_ids, d_dists, stats, nullptr, stream);
    return pgvrx_flat_exact_core(ix, d_queries, nq, k, row_begin, row_end, id_base, d_ids, d_dists, stats, stream);
}

}  // namespace

extern "C" {

int pgvrx_index_load_rows(pgvrx_index *ix, const void *rows, int64_t n) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (n < 0 || n > ix->desc.capacity) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "n exceeds the handle's capacity");
    if (n > 0 && !rows) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null rows");
    if (ix->build) return pgvrx_fail(ix, PGVRX_ERR_STATE, "handle holds a built graph");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    ix->generation++;
    ix->loaded = false;
    ix->rows_loaded = false;
    ix->tc_ready = false;
    CU_TRY(ix, ix->vectors.ensure(std::max<size_t>(16, (size_t)n * ix->row_stride)));
    if (n > 0) {
        if (ix->row_stride == ix->row_bytes) {
            CU_TRY(ix, cudaMemcpyAsync(ix->vectors.p, rows, (size_t)n * ix->row_bytes, cudaMemcpyHostToDevice, ix->stream));
        } else {
            CU_TRY(ix, cudaMemsetAsync(ix->vectors.p, 0, (size_t)n * ix->row_stride, ix->stream));
            CU_TRY(ix, cudaMemcpy2DAsync(ix->vectors.p, ix->row_stride, rows, ix->row_bytes, ix->row_bytes, (size_t)n,
                                         cudaMemcpyHostToDevice, ix->stream));
        }
        CU_TRY(ix, cudaStreamSynchronize(ix->stream));
    }
    ix->n = n;
    ix->n_upper = 0;
    ix->entry = -1;
    ix->entry_level = -1;
    ix->rows_loaded = true;
    return PGVRX_OK;
}

int pgvrx_flat_topk_device(pgvrx_index *ix, const void *d_queries, int64_t nq, int32_t k, int64_t row_begin, int64_t row_end,
                           int64_t id_base, int32_t *d_ids, float *d_dists, pgvrx_flat_stats *stats, void *cuda_stream) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (nq > 0 && (!d_queries || !d_ids || !d_dists)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null device buffer");
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
    return flat_core(ix, d_queries, nq, k, row_begin, row_end < 0 ? ix->n : row_end, id_base, d_ids, d_dists, stats, s);
}

int pgvrx_flat_topk(pgvrx_index *ix, const void *queries, int64_t nq, int32_t k, int32_t *out_ids, float *out_dists,
                    pgvrx_flat_stats *stats) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (nq > 0 && (!queries || !out_ids || !out_dists)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null buffer");
    if (nq < 0 || k < 1 || k > FLAT_MAX_K) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad k or nq");
    if (!ix->rows_loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no rows loaded");
    if (nq == 0) { if (stats) memset(stats, 0, sizeof *stats); return PGVRX_OK; }
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    CU_TRY(ix, ix->q.ensure((size_t)nq * ix->row_bytes));
    CU_TRY(ix, ix->ids.ensure((size_t)nq * k * 4));
    CU_TRY(ix, ix->dists.ensure((size_t)nq * k * 4));
    CU_TRY(ix, cudaMemcpyAsync(ix->q.p, queries, (size_t)nq * ix->row_bytes, cudaMemcpyHostToDevice, ix->stream));
    int rc = flat_core(ix, ix->q.p, nq, k, 0, ix->n, 0, (int32_t *)ix->ids.p, (float *)ix->dists.p, stats, ix->stream);
    if (rc) return rc;
    CU_TRY(ix, cudaMemcpyAsync(out_ids, ix->ids.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    CU_TRY(ix, cudaMemcpyAsync(out_dists, ix->dists.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    CU_TRY(ix, cudaStreamSynchronize(ix->stream));
    return PGVRX_OK;
}

int pgvrx_rerank_device(pgvrx_index *ix, const void *d_queries, int64_t nq, const int32_t *d_cand, int32_t nc, int32_t k,
                        int64_t id_base, int32_t *d_ids, float *d_dists, void *cuda_stream) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (!ix->rows_loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no rows loaded");
    if (k < 1 || k > FLAT_MAX_K || nc < 1 || nq < 0) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad k, nc or nq");
    if (nq > 0 && (!d_queries || !d_cand || !d_ids || !d_dists)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null device buffer");
    if (nq == 0) return PGVRX_OK;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    RerankParams p;
    p.vectors = (const uint8_t *)ix->vectors.p;
    p.row_stride = ix->row_stride;
    p.stage_stride = ix->row_stride + (((ix->row_stride / 16) & 1u) ? 0u : 16u);
    p.R = (int)std::max<uint32_t>(1, std::min<uint32_t>(32, (17u * 1024u) / p.stage_stride));
    p.n_rows = ix->n;
    p.id_base = id_base;
    p.queries = (const uint8_t *)d_queries;
    p.q_stride = (uint32_t)ix->row_bytes;
    p.q_bytes = (uint32_t)ix->row_bytes;
    p.cand = d_cand;
    p.nq = (int)nq;
    p.nc = nc;
    p.k = k;
    p.out_ids = d_ids;
    p.out_dists = d_dists;
    p.flags = ix->has_flags ? (const uint8_t *)ix->flags.p : nullptr;
    const size_t smem = rerank_smem_bytes(p.row_stride, p.stage_stride, p.R, k);
    if (smem > 227 * 1024) return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "re-rank configuration exceeds shared memory");
    const int grid = (int)std::min<int64_t>((nq + 3) / 4, (int64_t)ix->sm_count * 3);
    CU_TRY(ix, launch_rerank(ix->desc.elem_type, ix->desc.metric, p, grid, smem, s));
    return PGVRX_OK;
}

int pgvrx_rerank(pgvrx_index *ix, const void *queries, int64_t nq, const int32_t *cand_ids, int32_t nc, int32_t k,
                 int32_t *out_ids, float *out_dists) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (nq > 0 && (!queries || !cand_ids || !out_ids || !out_dists)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null buffer");
    if (nq < 0 || nc < 1 || k < 1 || k > FLAT_MAX_K) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad k, nc or nq");
    if (!ix->rows_loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no rows loaded");
    if (nq == 0) return PGVRX_OK;
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    CU_TRY(ix, ix->q.ensure((size_t)nq * ix->row_bytes));
    CU_TRY(ix, ix->cand.ensure((size_t)nq * nc * 4));
    CU_TRY(ix, ix->ids.ensure((size_t)nq * k * 4));
    CU_TRY(ix, ix->dists.ensure((size_t)nq * k * 4));
    CU_TRY(ix, cudaMemcpyAsync(ix->q.p, queries, (size_t)nq * ix->row_bytes, cudaMemcpyHostToDevice, ix->stream));
    CU_TRY(ix, cudaMemcpyAsync(ix->cand.p, cand_ids, (size_t)nq * nc * 4, cudaMemcpyHostToDevice, ix->stream));
    int rc = pgvrx_rerank_device(ix, ix->q.p, nq, (const int32_t *)ix->cand.p, nc, k, 0, (int32_t *)ix->ids.p,
                                 (float *)ix->dists.p, ix->stream);
    if (rc) return rc;
    CU_TRY(ix, cudaMemcpyAsync(out_ids, ix->ids.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    CU_TRY(ix, cudaMemcpyAsync(out_dists, ix->dists.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    CU_TRY(ix, cudaStreamSynchronize(ix->stream));
    return PGVRX_OK;
}

int pgvrx_topk_merge_device(pgvrx_index *ix, const int32_t *d_part_ids, const float *d_part_dists, int32_t n_parts, int64_t nq,
                            int32_t k_in, int32_t k_out, int32_t *d_out_ids, float *d_out_dists, void *cuda_stream) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (n_parts < 1 || k_in < 1 || k_out < 1 || k_out > FLAT_MAX_K || nq < 0) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad merge shape");
    if (nq > 0 && (!d_part_ids || !d_part_dists || !d_out_ids || !d_out_dists)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null device buffer");
    if (nq == 0) return PGVRX_OK;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    CU_TRY(ix, launch_topk_merge(d_part_ids, d_part_dists, n_parts, (int)nq, k_in, k_out, d_out_ids, d_out_dists, s));
    return PGVRX_OK;
}

// The same merge over per-part pointers: the lists need not be contiguous with each other (one packed all-gather of
// (ids | distances) per rank leaves every rank's two lists at different strides; peers' lists read in place).
int pgvrx_topk_merge_parts_device(pgvrx_index *ix, const int32_t *const *d_part_ids, const float *const *d_part_dists, int32_t n_parts,
                                  int64_t nq, int32_t k_in, int32_t k_out, int32_t *d_out_ids, float *d_out_dists, void *cuda_stream) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (n_parts < 1 || n_parts > MERGE_MAX_PARTS || k_in < 1 || k_out < 1 || k_out > FLAT_MAX_K || nq < 0)
        return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad merge shape");
    if (nq > 0 && (!d_part_ids || !d_part_dists || !d_out_ids || !d_out_dists)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null device buffer");
    if (nq == 0) return PGVRX_OK;
    MergeParts parts;
    parts.n_parts = n_parts;
    for (int i = 0; i < n_parts; i++) {
        if (!d_part_ids[i] || !d_part_dists[i]) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null part");
        parts.ids[i] = d_part_ids[i];
        parts.dists[i] = d_part_dists[i];
    }
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    CU_TRY(ix, launch_topk_merge_parts(parts, (int)nq, k_in, k_out, d_out_ids, d_out_dists, s));
    return PGVRX_OK;
}

}  // extern "C"
