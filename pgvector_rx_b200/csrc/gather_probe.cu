// gather_probe.cu -- the ceiling the scan kernel's row gathers are measured against.
//
// K1 (hnsw_search.cu) fetches vector rows with one TMA bulk copy per row into a per-warp staging area.  Its roofline
// line quotes the measured HBM copy rate, but what bounds random row_stride-byte gathers that are mostly served from L2
// is the L2 -> SM fabric, not DRAM.  This kernel measures that ceiling on the index's own vector table with the scan's
// own transfer mechanism and geometry -- same persistent one-warp CTAs, same staging stride, same eviction policy, one
// mbarrier phase per batch of rows -- and nothing else: no visited set, no distances, no result list.  Row ids come from
// a pre-generated stream (untimed), either uniform over the table or a random walk over the layer-0 adjacency, which
// reproduces the hub reuse of real searches (and therefore their L2 hit rate).
#include <algorithm>
#include <cstring>

#include "hnsw_device.cuh"
#include "index_internal.h"

namespace pgvrx {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// ids[(w * steps + s) * 32 + lane]: the rows warp w gathers in step s
__global__ void gather_ids_kernel(GraphView g, int mode, int n_warps, int steps, int rows_per_step, uint32_t seed, int32_t *ids) {
    const int w = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n_warps) return;
    const int lm = 2 * g.m;
    uint32_t v = mix32(seed ^ (uint32_t)w * 0x9E3779B1u) % (uint32_t)g.n;
    for (int s = 0; s < steps; s++) {
        int32_t id;
        if (mode == 0 || g.adj0 == nullptr) {
            id = (int32_t)(mix32(seed + (uint32_t)w * 977u + (uint32_t)s * 131071u + (uint32_t)lane * 0x85EBCA6Bu) % (uint32_t)g.n);
        } else {
            // neighbour list of the current node (stored order, dense prefix); short lists are topped up with the node itself
            const int32_t nb = lane < lm ? __ldg(g.adj0 + (size_t)v * lm + lane) : -1;
            const uint32_t vm = __ballot_sync(FULL, nb >= 0);
            id = nb >= 0 ? nb : (int32_t)v;
            const int cnt = __popc(vm);
            const uint32_t pick = cnt > 0 ? mix32(seed ^ (uint32_t)(w * 31 + s)) % (uint32_t)cnt : 0u;
            const int32_t next = __shfl_sync(FULL, id, (int)pick);
            // restart now and then so that walks do not collapse into one region
            v = (s % 48 == 47) ? mix32(seed + (uint32_t)w + (uint32_t)s * 2654435761u) % (uint32_t)g.n : (uint32_t)next;
        }
        if (lane < rows_per_step) ids[((size_t)w * steps + s) * 32 + lane] = id;
    }
}

// one warp per CTA (the scan's geometry); `depth` batches of `R` rows in flight per warp
__global__ void __launch_bounds__(32) gather_probe_kernel(const uint8_t *__restrict__ vectors, uint32_t row_stride, uint32_t stage_stride,
                                                          const int32_t *__restrict__ ids, int steps, int R, int depth, int consume,
                                                          uint32_t *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x;
    const size_t slot_bytes = (size_t)R * stage_stride;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + (size_t)depth * slot_bytes);
    if (lane == 0) {
        for (int d = 0; d < depth; d++) mbar_init(mbar + d, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const uint64_t policy = l2_policy_evict_first();
    const int32_t *my = ids + (size_t)blockIdx.x * steps * 32;
    const int nvec = (int)(row_stride / 16);
    uint32_t acc = 0;
    uint32_t phase_bits = 0;                          // bit d = parity the next wait on slot d expects
    for (int s = 0; s < steps + depth - 1; s++) {
        if (s < steps) {
            const int d = s % depth;
            const int32_t id = lane < R ? __ldg(my + (size_t)s * 32 + lane) : 0;
            if (lane == 0) mbar_arrive_expect_tx(mbar + d, (uint32_t)R * row_stride);
            __syncwarp();
            if (lane < R)
                tma_row_g2s(smem + (size_t)d * slot_bytes + (size_t)lane * stage_stride, vectors + (size_t)id * row_stride, row_stride,
                            mbar + d, policy);
        }
        const int t = s - (depth - 1);
        if (t >= 0) {
            const int d = t % depth;
            mbar_wait(mbar + d, (phase_bits >> d) & 1u);
            phase_bits ^= 1u << d;
            if (lane < R) {
                const uint4 *r4 = reinterpret_cast<const uint4 *>(smem + (size_t)d * slot_bytes + (size_t)lane * stage_stride);
                if (consume) { for (int j = 0; j < nvec; j++) { const uint4 x = r4[j]; acc += x.x ^ x.y ^ x.z ^ x.w; } }
                else { const uint4 x = r4[0]; acc += x.x; }
            }
            __syncwarp();                             // all lanes done with the slot before it is refilled
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;            // keeps the loads alive
}

}  // namespace pgvrx

using namespace pgvrx;

extern "C" int pgvrx_gather_probe(pgvrx_index *ix, int32_t mode, int32_t warps_per_sm, int32_t rows_per_step, int32_t depth,
                                  int32_t steps, int32_t consume, pgvrx_gather_probe_result *out) {
    if (!ix || !out) return PGVRX_ERR_INVALID;
    memset(out, 0, sizeof *out);
    if (!ix->rows_loaded || ix->n < 1) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no rows loaded");
    if (mode < 0 || mode > 1 || rows_per_step < 1 || rows_per_step > 32 || depth < 1 || depth > 8 || steps < 1 || steps > (1 << 16))
        return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad gather probe geometry");
    if (mode == 1 && !ix->loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "the random-walk id stream needs a loaded graph");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    const uint32_t stage_stride = ix->row_stride + (((ix->row_stride / 16) & 1u) ? 0u : 16u);   // the scan's stride (capi.cu)
    const size_t smem = (size_t)depth * rows_per_step * stage_stride + 64;
    if (smem > 227 * 1024) return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "gather probe geometry exceeds shared memory");
    CU_TRY(ix, cudaFuncSetAttribute(gather_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CU_TRY(ix, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gather_probe_kernel, 32, smem));
    if (occ < 1) return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "gather probe does not fit on an SM");
    const int wps = warps_per_sm > 0 ? std::min<int>(warps_per_sm, occ) : occ;
    const int grid = wps * ix->sm_count;
    DevBuf ids, sink;
    CU_TRY(ix, ids.ensure((size_t)grid * steps * 32 * 4));
    CU_TRY(ix, sink.ensure(64));
    cudaStream_t s = ix->stream;
    CU_TRY(ix, cudaMemsetAsync(ids.p, 0, (size_t)grid * steps * 32 * 4, s));
    GraphView g = pgvrx_view_of(ix);
    if (!ix->loaded) g.adj0 = nullptr;
    gather_ids_kernel<<<(grid + 3) / 4, 128, 0, s>>>(g, mode, grid, steps, rows_per_step, 12345u, (int32_t *)ids.p);
    CU_TRY(ix, cudaGetLastError());
    auto launch = [&]() {
        gather_probe_kernel<<<grid, 32, smem, s>>>((const uint8_t *)ix->vectors.p, ix->row_stride, stage_stride, (const int32_t *)ids.p,
                                                   steps, rows_per_step, depth, consume, (uint32_t *)sink.p);
    };
    launch();                                          // warm-up (also warms L2 the way a steady stream of batches does)
    CU_TRY(ix, cudaGetLastError());
    const int reps = 3;
    CU_TRY(ix, cudaEventRecord(ix->ev0, s));
    for (int r = 0; r < reps; r++) launch();
    CU_TRY(ix, cudaEventRecord(ix->ev1, s));
    cudaError_t e = cudaStreamSynchronize(s);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ix->ev0, ix->ev1);
    ids.release();
    sink.release();
    CU_TRY(ix, e);
    out->bytes_per_launch = (uint64_t)grid * (uint64_t)steps * (uint64_t)rows_per_step * ix->row_stride;
    out->ms_per_launch = ms / reps;
    out->gb_per_s = (double)out->bytes_per_launch / (out->ms_per_launch * 1e-3) / 1e9;
    out->warps_per_sm = wps;
    out->grid = grid;
    out->row_bytes = ix->row_stride;
    return PGVRX_OK;
}
