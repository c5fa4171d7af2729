// hnsw_search.cu -- K1: batched multi-layer HNSW scan (get_scan_items, src/index/scan.rs:458-530).
//
// Persistent grid, one warp per query, queries handed out through an atomic counter so uneven
// search lengths balance.  See hnsw_device.cuh for the per-warp machinery and DESIGN.md for the
// data layout and the roofline that bounds this kernel (HBM, gathered vector rows).
#include "hnsw_device.cuh"
#include "hnsw_kernels.h"

namespace pgvrx {

__host__ __device__ inline uint32_t round16(uint32_t x) { return (x + 15u) & ~15u; }

// shared memory bytes one warp needs
size_t search_warp_smem(const SearchLaunch &L) {
    size_t b = L.stage_bytes;                        // staging rows
    b += L.row_stride;                               // query
    if (L.list_j == 0) b += round16((uint32_t)((L.ef + L.tail_cap + 1) * 8));  // S (else it lives in registers)
    b += 128;                                        // nbr ids
    b += 16;                                         // mbarrier
    return b;
}

template <int J> struct ListOf {
    typedef RegList<J> type;
    __device__ static __forceinline__ void attach(type &, uint8_t *) {}
};
template <> struct ListOf<0> {
    typedef SmemList type;
    __device__ static __forceinline__ void attach(type &l, uint8_t *mem) { l.S = reinterpret_cast<uint64_t *>(mem); l.hint = 0; }
};

// LISTJ: 0 = result list in shared memory, else registers per lane (list capacity 32 * LISTJ)
// FLAGGED: the index carries per-element vacuum flags (GraphView::flags != NULL); only built for the shared-memory list
template <int ELEM, int METRIC, int LISTJ, bool FLAGGED>
__global__ void __launch_bounds__(SEARCH_THREADS) hnsw_scan_kernel(GraphView g, SearchParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cap = p.ef + p.tail_cap;
    const size_t s_bytes = LISTJ == 0 ? round16((uint32_t)((cap + 1) * 8)) : 0;
    const size_t per_warp = (size_t)p.stage_bytes + g.row_stride + s_bytes + 128 + 16;
    uint8_t *base = smem + (size_t)warp * per_warp;

    WarpCtx c;
    c.stage = base;
    c.qsm = c.stage + (size_t)p.stage_bytes;
    uint8_t *s_mem = c.qsm + g.row_stride;
    c.nbr = reinterpret_cast<int32_t *>(s_mem + s_bytes);
    c.mbar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(c.nbr) + 128);
    typename ListOf<LISTJ>::type S;
    ListOf<LISTJ>::attach(S, s_mem);
    c.stage_stride = p.stage_stride;
    c.R = p.R;
    c.chunk_bytes = p.chunk_bytes;
    c.chunk_stride = p.chunk_stride;
    c.nvec = (int)(g.row_stride / 16);
    c.cap = cap;
    c.lane = lane;
    c.phase = 0;
    c.phase2 = 0;
    c.stage2 = (c.chunk_bytes != 0 && (uint32_t)p.stage_bytes >= 64u * c.chunk_stride) ? c.stage + 32u * (size_t)c.chunk_stride : nullptr;
    c.policy = l2_policy_evict_first();
    c.rs_log = nullptr;
    c.rs_cap = c.rs_n = 0;
    c.rs_ent = 3;
    c.spec = p.spec == 1;
    c.hybrid = p.spec == 2 && p.chunk_bytes == 0 && p.R < 32;   // (one round holds every list: plain speculative order)
    if (p.spec == 2 && !c.hybrid) c.spec = true;
    const int slot = blockIdx.x * (blockDim.x / 32) + warp;
    c.vis = p.vis_base + (size_t)slot * p.vis_entries;

    if (lane == 0) {
        mbar_init(c.mbar, 1);
        mbar_init(c.mbar + 1, 1);
        fence_mbar_init();
    }
    __syncwarp();

    const uint32_t up_entries = p.up_entries;

    while (true) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.work_counter, 1);
        w = __shfl_sync(FULL, w, 0);
        const int nwork = p.qmap ? *p.n_work : p.nq;
        if (w >= nwork) break;
        const int qi = p.qmap ? p.qmap[w] : w;

        c.n_dist = c.n_exp0 = c.n_expu = c.n_vis = 0;
        c.status = 0;
#ifdef PGVRX_PHASE_TIMING
        for (int t = 0; t < 8; t++) c.t[t] = 0;
        const long long t_query0 = clock64();
#endif

        // ---- query row -> shared memory (zero padded to row_stride) ----
        {
            const uint8_t *q = p.queries + (size_t)qi * p.q_stride;
            if ((p.q_stride & 15u) == 0 && p.q_bytes == g.row_stride) {
                const uint4 *q4 = reinterpret_cast<const uint4 *>(q);
                uint4 *d4 = reinterpret_cast<uint4 *>(c.qsm);
                for (int i = lane; i < c.nvec; i += 32) d4[i] = __ldg(q4 + i);
            } else {
                for (uint32_t i = lane; i < g.row_stride; i += 32) c.qsm[i] = i < p.q_bytes ? __ldg(q + i) : (uint8_t)0;
            }
        }
        __syncwarp();

        int out_n = 0, w_n = 0;
        // empty index => no rows (scan.rs:470-472); so does an entry point flagged deleted (scan.rs:475-486 -> :179)
        bool entry_ok = g.entry >= 0;
        if constexpr (FLAGGED) entry_ok = entry_ok && !(__ldg(g.flags + g.entry) & ELEM_FLAG_DELETED);
        if (entry_ok) {
            // ---- distance to the entry point (scan.rs:475-486) ----
            if (lane == 0) c.nbr[0] = g.entry;
            __syncwarp();
            gather_rows(c, g, 0, 1);
            float d0 = 0.0f;
            if (lane == 0) d0 = row_distance<ELEM, METRIC>(c.qsm, c.stage, c.nvec);
            d0 = __shfl_sync(FULL, d0, 0);
            c.n_dist = 1;
            __syncwarp();
            uint64_t ep_key = make_key(d0, (uint32_t)g.entry);

            // ---- greedy descent: ef = 1, fresh visited set per layer (scan.rs:492-512) ----
            c.ef = 1;
            for (int lc = g.entry_level; lc >= 1 && c.status == 0; lc--) {
                vis_clear(c.vis, up_entries, lane);
                if (lane == 0) vis_insert(c.vis, up_entries - 1, key_id(ep_key));
                __syncwarp();
                c.n_vis++;
                S.reset1(c, ep_key);
                search_layer<ELEM, METRIC, FLAGGED>(c, S, g, lc, up_entries - 1, up_entries / 4 * 3 - 32);
                ep_key = S.get(c, 0) & ~1ull;     // nearest of W becomes the next entry point (scan.rs:510)
            }
            // ---- layer 0 with ef_search (scan.rs:515-528) ----
            if (c.status == 0) {
                vis_clear(c.vis, p.vis_entries, lane);
                if (lane == 0) vis_insert(c.vis, p.vis_entries - 1, key_id(ep_key));
                __syncwarp();
                c.n_vis++;
                c.ef = p.ef;
                S.reset1(c, ep_key);
                search_layer<ELEM, METRIC, FLAGGED>(c, S, g, 0, p.vis_entries - 1, p.vis_entries / 4 * 3 - 32);
            }
            w_n = min(c.len, c.ef);               // W = S[0, ef); nearest first (scan.rs:861)
            out_n = min(w_n, p.k);
        }

        // ---- results ----
        if (c.status == 0) {
            if constexpr (!FLAGGED) {
                S.store(c, p.out_ids + (size_t)qi * p.k, p.out_dists + (size_t)qi * p.k, out_n);
            } else {
                // amgettuple skips elements without heap TIDs (scan.rs:866-868): the first k of W that have one
                int cnt = 0;
                for (int i = 0; i < w_n && cnt < p.k; i++) {
                    const uint64_t key = S.get(c, i);
                    if (__ldg(g.flags + key_id(key)) & ELEM_FLAG_NO_HEAPTIDS) continue;
                    if (lane == 0) {
                        p.out_ids[(size_t)qi * p.k + cnt] = (int32_t)key_id(key);
                        p.out_dists[(size_t)qi * p.k + cnt] = ord_f32(key_ord(key));
                    }
                    cnt++;
                }
                out_n = cnt;
            }
            for (int i = out_n + lane; i < p.k; i += 32) {
                p.out_ids[(size_t)qi * p.k + i] = -1;
                p.out_dists[(size_t)qi * p.k + i] = __int_as_float(0x7f800000);
            }
            if (lane == 0) p.out_counts[qi] = out_n;
        } else if (lane == 0) {
            p.out_counts[qi] = -1;
        }
        if (lane == 0) {
            p.out_status[qi] = c.status;
            if (p.out_ctr) {
                uint4 v = make_uint4(c.n_dist, c.n_exp0, c.n_expu, c.n_vis);
                reinterpret_cast<uint4 *>(p.out_ctr)[qi] = v;
            }
#ifdef PGVRX_PHASE_TIMING
            if (p.out_phase) {
                c.t[5] = (uint32_t)(clock64() - t_query0);
                for (int t = 0; t < 8; t++) p.out_phase[(size_t)qi * 8 + t] = c.t[t];
            }
#endif
        }
        __syncwarp();
    }
}

// gather the queries whose status is non-zero into qmap / n_work for the retry launch
__global__ void collect_failed_kernel(const uint32_t *status, int nq, int32_t *qmap, int *n_work) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq && status[i] != 0) qmap[atomicAdd(n_work, 1)] = i;
}

typedef void (*scan_kernel_t)(GraphView, SearchParams);

template <int ELEM, int METRIC>
static scan_kernel_t pick_list(int list_j, bool flagged) {
    if (flagged) return list_j == 0 ? hnsw_scan_kernel<ELEM, METRIC, 0, true> : nullptr;
    switch (list_j) {
        case 2: return hnsw_scan_kernel<ELEM, METRIC, 2, false>;
        case 4: return hnsw_scan_kernel<ELEM, METRIC, 4, false>;
        case 8: return hnsw_scan_kernel<ELEM, METRIC, 8, false>;
        case 0: return hnsw_scan_kernel<ELEM, METRIC, 0, false>;
    }
    return nullptr;
}

static scan_kernel_t pick_kernel(int elem, int metric, int list_j, bool flagged) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return pick_list<ELEM_F32, METRIC_L2SQ>(list_j, flagged);
        if (metric == METRIC_NEG_IP) return pick_list<ELEM_F32, METRIC_NEG_IP>(list_j, flagged);
        if (metric == METRIC_L1) return pick_list<ELEM_F32, METRIC_L1>(list_j, flagged);
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return pick_list<ELEM_F16, METRIC_L2SQ>(list_j, flagged);
        if (metric == METRIC_NEG_IP) return pick_list<ELEM_F16, METRIC_NEG_IP>(list_j, flagged);
        if (metric == METRIC_L1) return pick_list<ELEM_F16, METRIC_L1>(list_j, flagged);
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return pick_list<ELEM_BIT, METRIC_HAMMING>(list_j, flagged);
        if (metric == METRIC_JACCARD) return pick_list<ELEM_BIT, METRIC_JACCARD>(list_j, flagged);
    }
    return nullptr;
}

cudaError_t scan_kernel_occupancy(int elem, int metric, int list_j, bool flagged, size_t smem_block, int threads, int *blocks_per_sm) {
    scan_kernel_t k = pick_kernel(elem, metric, list_j, flagged);
    if (!k) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_block);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads, smem_block);
}

cudaError_t launch_scan_kernel(int elem, int metric, int list_j, const GraphView &g, const SearchParams &p, int grid,
                               size_t smem_block, int threads, cudaStream_t stream) {
    scan_kernel_t k = pick_kernel(elem, metric, list_j, g.flags != nullptr);
    if (!k) return cudaErrorInvalidValue;
    k<<<grid, threads, smem_block, stream>>>(g, p);
    return cudaGetLastError();
}

cudaError_t launch_collect_failed(const uint32_t *status, int nq, int32_t *qmap, int *n_work, cudaStream_t stream) {
    collect_failed_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(status, nq, qmap, n_work);
    return cudaGetLastError();
}

}  // namespace pgvrx
