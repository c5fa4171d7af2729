// hnsw_build.cu -- K2: the search half of find_element_neighbors (src/graph/mod.rs:355-427) for a
// batch of pending inserts, each run by one warp against the same committed snapshot of the graph.
//
// Per insert: distance to the entry point (:371-377), greedy ef = 1 descent through layers
// entry_level .. level+1 (:385-399), then search_layer at ef_construction on layers
// min(level, entry_level) .. 0 with W of one layer as the entry points of the next (:403-426).
// Outputs per searched layer: W sorted ascending by (distance, id) -- what select_neighbors consumes
// on the host -- and the ids of all expanded nodes (the search's read set), which lets the host
// prove that a speculative result is still exact after earlier inserts of the batch were committed.
#include <algorithm>

#include "hnsw_device.cuh"
#include "hnsw_kernels.h"

namespace pgvrx {

__host__ __device__ inline uint32_t round16b(uint32_t x) { return (x + 15u) & ~15u; }

template <int J> struct BListOf {
    typedef RegList<J> type;
    __device__ static __forceinline__ void attach(type &, uint8_t *) {}
};
template <> struct BListOf<0> {
    typedef SmemList type;
    __device__ static __forceinline__ void attach(type &l, uint8_t *mem) { l.S = reinterpret_cast<uint64_t *>(mem); l.hint = 0; }
};

// FLAGGED: the index carries vacuum flags (inserts into a vacuumed index): elements flagged deleted are visited and nothing
// else (load_element, scan.rs:179), and a deleted entry point leaves the new element without neighbours
// (find_element_neighbors_on_disk, insert.rs:1036-1047).  Only built for the shared-memory list.
template <int ELEM, int METRIC, int LISTJ, bool FLAGGED = false>
__global__ void __launch_bounds__(SEARCH_THREADS) hnsw_build_search_kernel(GraphView g, BuildParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cap = p.efc + p.tail_cap;
    const size_t s_bytes = LISTJ == 0 ? round16b((uint32_t)((cap + 1) * 8)) : 0;
    const size_t per_warp = (size_t)p.stage_bytes + g.row_stride + s_bytes + 128 + 16;
    uint8_t *base = smem + (size_t)warp * per_warp;

    WarpCtx c;
    c.stage = base;
    c.qsm = c.stage + (size_t)p.stage_bytes;
    uint8_t *s_mem = c.qsm + g.row_stride;
    c.nbr = reinterpret_cast<int32_t *>(s_mem + s_bytes);
    c.mbar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(c.nbr) + 128);
    typename BListOf<LISTJ>::type S;
    BListOf<LISTJ>::attach(S, s_mem);
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
        if (w >= p.n_pending) break;
        const int e = p.elem_ids[w];
        const int lvl = g.level[e];

        c.n_dist = c.n_exp0 = c.n_expu = c.n_vis = 0;
        c.status = 0;
        c.rs_log = p.rs_log + (size_t)w * p.rs_cap * p.rs_ent;
        c.rs_ent = p.rs_ent;
        c.spec = true;                 // the log needs the distance of every entry of every list read
        c.hybrid = false;
        c.rs_cap = p.rs_cap;
        c.rs_n = 0;
        for (int l = lane; l < p.max_layers; l += 32) p.w_counts[(size_t)w * p.max_layers + l] = -1;
        if constexpr (FLAGGED) {
            if (__ldg(g.flags + g.entry) & ELEM_FLAG_DELETED) {       // load_element(entry) => None: every list stays empty
                __syncwarp();
                for (int l = lane; l < p.max_layers && l <= min(lvl, g.entry_level); l += 32) p.w_counts[(size_t)w * p.max_layers + l] = 0;
                if (lane == 0) {
                    p.out_status[w] = 0;
                    p.rs_counts[w] = 0;
                    if (p.out_ctr) reinterpret_cast<uint4 *>(p.out_ctr)[w] = make_uint4(0, 0, 0, 0);
                }
                __syncwarp();
                continue;
            }
        }

        // the new element's own row is the query (build.rs:444-454 copies it into the arena first)
        {
            const uint4 *q4 = reinterpret_cast<const uint4 *>(g.vectors + (size_t)e * g.row_stride);
            uint4 *d4 = reinterpret_cast<uint4 *>(c.qsm);
            for (int i = lane; i < c.nvec; i += 32) d4[i] = __ldg(q4 + i);
        }
        if (lane == 0) c.nbr[0] = g.entry;
        __syncwarp();
        gather_rows(c, g, 0, 1);
        float d0 = 0.0f;
        if (lane == 0) d0 = row_distance<ELEM, METRIC>(c.qsm, c.stage, c.nvec);    // graph/mod.rs:371-377
        d0 = __shfl_sync(FULL, d0, 0);
        c.n_dist = 1;
        __syncwarp();
        uint64_t ep_key = make_key(d0, (uint32_t)g.entry);

        // phase 1: greedy descent above the new element's level (graph/mod.rs:385-399)
        c.ef = 1;
        for (int lc = g.entry_level; lc > lvl && c.status == 0; lc--) {
            vis_clear(c.vis, up_entries, lane);
            if (lane == 0) vis_insert(c.vis, up_entries - 1, key_id(ep_key));
            __syncwarp();
            c.n_vis++;
            S.reset1(c, ep_key);
            search_layer<ELEM, METRIC, FLAGGED>(c, S, g, lc, up_entries - 1, up_entries / 4 * 3 - 32);
            ep_key = S.get(c, 0) & ~1ull;                                           // :396-398 w[0]
        }
        // phase 2: ef_construction search on every layer the new element joins (graph/mod.rs:403-426)
        c.ef = p.efc;
        bool first = true;
        for (int lc = min(lvl, g.entry_level); lc >= 0 && c.status == 0; lc--) {
            vis_clear(c.vis, p.vis_entries, lane);
            if (first) {
                if (lane == 0) vis_insert(c.vis, p.vis_entries - 1, key_id(ep_key));
                __syncwarp();
                c.n_vis++;
                S.reset1(c, ep_key);
            } else {
                S.reopen(c, c.vis, p.vis_entries - 1);                              // :425 ep = w
                c.n_vis += c.len;
            }
            first = false;
            search_layer<ELEM, METRIC, FLAGGED>(c, S, g, lc, p.vis_entries - 1, p.vis_entries / 4 * 3 - 32);
            if (c.status == 0 && lc < p.max_layers) {
                const int out_n = min(c.len, c.ef);
                const size_t o = ((size_t)w * p.max_layers + lc) * p.efc;
                S.store(c, p.w_ids + o, p.w_dists + o, out_n);
                if (lane == 0) p.w_counts[(size_t)w * p.max_layers + lc] = out_n;
            }
        }
        if (lane == 0) {
            p.out_status[w] = c.status;
            p.rs_counts[w] = c.rs_n;
            if (p.out_ctr) reinterpret_cast<uint4 *>(p.out_ctr)[w] = make_uint4(c.n_dist, c.n_exp0, c.n_expu, c.n_vis);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// Vacuum repair search: find_element_neighbors_on_disk(..., skip_set) (insert.rs:1021-1123) for elements that are
// already in the graph (vacuum.rs:288-407).  Same walk as the kernel above with three differences: members of the skip
// set (elements without heap TIDs, and the element itself) are traversed but do not count towards ef (search_layer's
// SKIP mode); ef = ef_construction + 1; elements flagged deleted by an EARLIER vacuum are visited and nothing else
// (FLAGGED).  W is written out whole, skip members included: the host filters them and keeps the nearest lm.
// Shared-memory list only.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool repair_is_skip(const GraphView &g, const WarpCtx &c, uint32_t id) {
    return (int32_t)id == c.self_id || (__ldg(g.flags + id) & ELEM_FLAG_NO_HEAPTIDS);
}

// seed a layer's search with one entry point (insert.rs:1050, :1069-1071)
__device__ __forceinline__ void repair_seed1(WarpCtx &c, SmemList &S, const GraphView &g, uint64_t ep_key, int ef) {
    c.ef = ef + (repair_is_skip(g, c, key_id(ep_key)) ? 1 : 0);
    c.excess = 0;
    S.reset1(c, ep_key);
    if (c.len >= c.ef) c.f_ord = key_ord(ep_key);
}

// W of the layer above becomes the entry points of this one (insert.rs:1119 ep_list = w): every member is pushed, the
// counted ones raise w_len -- possibly past ef, nothing is evicted while seeding (scan.rs:326-336)
__device__ __forceinline__ void repair_reopen(WarpCtx &c, SmemList &S, const GraphView &g, uint32_t *vis, uint32_t vis_mask, int ef) {
    c.len = min(c.len, c.ef);
    __syncwarp();
    int n_skip = 0;
    for (int base = 0; base < c.len; base += 32) {
        const int i = base + c.lane;
        bool sk = false;
        if (i < c.len) {
            S.S[i] &= ~1ull;
            vis_insert(vis, vis_mask, key_id(S.S[i]));
            sk = repair_is_skip(g, c, key_id(S.S[i]));
        }
        n_skip += __popc(__ballot_sync(FULL, sk));
    }
    __syncwarp();
    S.hint = 0;
    c.ef = max(ef, c.len - n_skip) + n_skip;
    c.excess = max(0, c.len - n_skip - ef);
    if (c.ef >= c.cap) { c.status |= ST_TIE_OVERFLOW; c.ef = c.cap - 1; }
    if (c.len >= c.ef) c.f_ord = key_ord(S.S[c.ef - 1]);
}

template <int ELEM, int METRIC>
__global__ void __launch_bounds__(SEARCH_THREADS) hnsw_repair_search_kernel(GraphView g, RepairParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cap = p.ef + p.tail_cap;
    const size_t s_bytes = round16b((uint32_t)((cap + 1) * 8));
    const size_t per_warp = (size_t)p.stage_bytes + g.row_stride + s_bytes + 128 + 16;
    uint8_t *base = smem + (size_t)warp * per_warp;

    WarpCtx c;
    c.stage = base;
    c.qsm = c.stage + (size_t)p.stage_bytes;
    uint8_t *s_mem = c.qsm + g.row_stride;
    c.nbr = reinterpret_cast<int32_t *>(s_mem + s_bytes);
    c.mbar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(c.nbr) + 128);
    SmemList S;
    S.S = reinterpret_cast<uint64_t *>(s_mem);
    S.hint = 0;
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
        if (w >= p.n_pending) break;
        const int e = p.elem_ids[w];
        const int lvl = g.level[e];

        c.n_dist = c.n_exp0 = c.n_expu = c.n_vis = 0;
        c.status = 0;
        c.rs_log = p.rs_log + (size_t)w * p.rs_cap * p.rs_ent;
        c.rs_ent = p.rs_ent;
        c.spec = true;
        c.hybrid = false;
        c.rs_cap = p.rs_cap;
        c.rs_n = 0;
        c.self_id = e;
        for (int l = lane; l < p.max_layers; l += 32) p.w_counts[(size_t)w * p.max_layers + l] = -1;
        {
            const uint4 *q4 = reinterpret_cast<const uint4 *>(g.vectors + (size_t)e * g.row_stride);
            uint4 *d4 = reinterpret_cast<uint4 *>(c.qsm);
            for (int i = lane; i < c.nvec; i += 32) d4[i] = __ldg(q4 + i);
        }
        __syncwarp();
        // load_element(entry) fails for an entry point flagged deleted: every list comes back empty (insert.rs:1036-1047)
        const bool entry_ok = g.entry >= 0 && !(__ldg(g.flags + g.entry) & ELEM_FLAG_DELETED);
        if (entry_ok) {
            if (lane == 0) c.nbr[0] = g.entry;
            __syncwarp();
            gather_rows(c, g, 0, 1);
            float d0 = 0.0f;
            if (lane == 0) d0 = row_distance<ELEM, METRIC>(c.qsm, c.stage, c.nvec);
            d0 = __shfl_sync(FULL, d0, 0);
            c.n_dist = 1;
            __syncwarp();
            uint64_t ep_key = make_key(d0, (uint32_t)g.entry);

            // phase 1: ef = 1 above the element's level (insert.rs:1052-1072)
            for (int lc = g.entry_level; lc > lvl && c.status == 0; lc--) {
                vis_clear(c.vis, up_entries, lane);
                if (lane == 0) vis_insert(c.vis, up_entries - 1, key_id(ep_key));
                __syncwarp();
                c.n_vis++;
                repair_seed1(c, S, g, ep_key, 1);
                search_layer<ELEM, METRIC, true, true>(c, S, g, lc, up_entries - 1, up_entries / 4 * 3 - 32);
                ep_key = S.get(c, 0) & ~1ull;                                       // nearest of W, skip member or not
            }
            // phase 2: ef_construction + 1 on the element's layers (insert.rs:1075-1120)
            bool first = true;
            for (int lc = min(lvl, g.entry_level); lc >= 0 && c.status == 0; lc--) {
                vis_clear(c.vis, p.vis_entries, lane);
                if (first) {
                    if (lane == 0) vis_insert(c.vis, p.vis_entries - 1, key_id(ep_key));
                    __syncwarp();
                    c.n_vis++;
                    repair_seed1(c, S, g, ep_key, p.ef);
                } else {
                    repair_reopen(c, S, g, c.vis, p.vis_entries - 1, p.ef);
                    c.n_vis += c.len;
                }
                first = false;
                search_layer<ELEM, METRIC, true, true>(c, S, g, lc, p.vis_entries - 1, p.vis_entries / 4 * 3 - 32);
                if (c.status == 0 && lc < p.max_layers) {
                    const int out_n = min(c.len, c.ef);
                    if (out_n > p.w_pitch) c.status |= ST_TIE_OVERFLOW;            // more skip members than the output row holds
                    else {
                        const size_t o = ((size_t)w * p.max_layers + lc) * p.w_pitch;
                        S.store(c, p.w_ids + o, p.w_dists + o, out_n);
                        if (lane == 0) p.w_counts[(size_t)w * p.max_layers + lc] = out_n;
                    }
                }
            }
        }
        if (lane == 0) {
            p.out_status[w] = c.status;
            p.rs_counts[w] = c.rs_n;
            if (p.out_ctr) reinterpret_cast<uint4 *>(p.out_ctr)[w] = make_uint4(c.n_dist, c.n_exp0, c.n_expu, c.n_vis);
        }
        __syncwarp();
    }
}

// scatter freshly committed neighbour rows into the device adjacency (one thread per int)
__global__ void scatter_rows_kernel(int32_t *dst, const int64_t *row_offsets, const int32_t *packed, int row_len, int n_rows) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int r = t / row_len, j = t - r * row_len;
    if (r < n_rows) dst[row_offsets[r] + j] = packed[(size_t)r * row_len + j];
}


// ---------------------------------------------------------------------------------------------------
// check_element_closer for whole groups (graph/mod.rs:315-339), used when update_neighbor_connections
// re-selects a full list (:473-485).  A group is a node's neighbour list plus the pending elements that link
// back to it; every member a carries its distance to the node, own[a].  The heuristic only ever asks
// "is d(a, r) <= own[a]" for some already kept r, so the kernel answers that for ALL pairs at once:
//     mask[a] bit b  =  d(a, b) <= own[a]
// (128 bits per member).  On the host a candidate is then rejected iff mask[a] & kept != 0.
// One CTA per group: the rows are staged once (one TMA bulk copy per row), lane <-> column b, row a is
// broadcast from shared memory; distances use the same sequential arithmetic as everywhere else.
// ---------------------------------------------------------------------------------------------------
// Every distance here is symmetric bit for bit (the element-wise terms commute exactly and are added in the same element
// order), so each unordered pair is evaluated once and feeds both masks: bit b of mask[a] and bit a of mask[b].  Sixteen
// warps share a group: a group of 128 wide rows is 16 k chains of up to thousands of dependent additions each, and four
// warps -- one per scheduler -- left every chain's latency exposed (C3: 0.46 ms per group, more than the searches of the
// round).
constexpr int PAIR_THREADS = 512;
template <int ELEM, int METRIC>
__global__ void __launch_bounds__(PAIR_THREADS) pairwise_kernel(GraphView g, PairParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    uint8_t *stage = smem + 128;
    constexpr int MW = PAIR_MAX_ROWS / 32;
    __shared__ int s_group;
    __shared__ uint32_t s_mask[PAIR_MAX_ROWS * MW];
    __shared__ float s_own[PAIR_MAX_ROWS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    __syncthreads();
    const uint64_t policy = l2_policy_evict_first();
    const int nvec = (int)(g.row_stride / 16);
    uint32_t phase = 0;
    while (true) {
        if (t == 0) s_group = atomicAdd(p.work_counter, 1);
        __syncthreads();
        const int gi = s_group;
        if (gi >= p.n_groups) break;
        // explicit groups (back-link lists) or, in W mode, group gi = the candidate list W of (pending element, layer)
        // as the search kernel left it: fixed pitch, length in w_counts, own distance = distance to the new element
        int off, cnt;
        if (p.group_off) {
            off = p.group_off[gi];
            cnt = p.group_off[gi + 1] - off;
        } else {
            off = gi * p.w_pitch;
            cnt = p.w_counts[gi];
            const int lm = (gi % p.w_layers) == 0 ? 2 * g.m : g.m;
            if (cnt <= lm) cnt = 0;                    // select_neighbors returns such a list unchanged (graph/mod.rs:276)
        }
        if (cnt == 0) { __syncthreads(); continue; }
        if (t == 0) mbar_arrive_expect_tx(mbar, (uint32_t)cnt * g.row_stride);
        for (int i = t; i < PAIR_MAX_ROWS * MW; i += PAIR_THREADS) s_mask[i] = 0u;
        if (t < cnt) s_own[t] = p.own_dist[off + t];
        __syncthreads();
        if (t < cnt)
            tma_row_g2s(stage + (size_t)t * p.stage_stride, g.vectors + (size_t)p.ids[off + t] * g.row_stride, g.row_stride, mbar, policy);
        mbar_wait(mbar, phase);
        phase ^= 1u;
        // work items: (row a, block of 32 columns) for the blocks that reach column a or beyond; lanes on columns b >= a
        const int nblk = (cnt + 31) >> 5;
        for (int item = warp; item < cnt * nblk; item += PAIR_THREADS / 32) {
            const int a = item / nblk, b0 = (item - a * nblk) << 5;
            if (b0 + 31 < a) continue;
            const int b = b0 + lane;
            const bool act = b >= a && b < cnt;
            float d = 0.0f;
            if (act) d = row_distance<ELEM, METRIC>(stage + (size_t)a * p.stage_stride, stage + (size_t)b * p.stage_stride, nvec);
            const uint32_t wa = __ballot_sync(FULL, act && d <= s_own[a]);                    // :333, e = a against r = b
            if (lane == 0 && wa) atomicOr(&s_mask[a * MW + (b0 >> 5)], wa);
            if (act && b > a && d <= s_own[b]) atomicOr(&s_mask[b * MW + (a >> 5)], 1u << (a & 31));   // e = b against r = a
        }
        __syncthreads();
        for (int i = t; i < cnt * MW; i += PAIR_THREADS) p.out[(size_t)off * MW + i] = s_mask[i];
        // generic-proxy reads of the staging rows are ordered before the next group's TMA writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
}

// Distances between the pending elements of a round: out[j * n + i] = d(pending j, pending i) for i < j, evaluated as
// d(query = j, row = i) like every other distance.  The host needs them when it validates element j against a list
// that an earlier element i of the same round has joined (build_host.cu).  One warp per j, lanes over i.
template <int ELEM, int METRIC>
__global__ void __launch_bounds__(SEARCH_THREADS) pending_pairs_kernel(GraphView g, const int32_t *__restrict__ ids, int n,
                                                                       float *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *q = smem + (size_t)warp * g.row_stride;
    const int nvec = (int)(g.row_stride / 16);
    for (int j = blockIdx.x * (SEARCH_THREADS / 32) + warp; j < n; j += gridDim.x * (SEARCH_THREADS / 32)) {
        const uint4 *src = reinterpret_cast<const uint4 *>(g.vectors + (size_t)ids[j] * g.row_stride);
        __syncwarp();
        for (int v = lane; v < nvec; v += 32) reinterpret_cast<uint4 *>(q)[v] = __ldg(src + v);
        __syncwarp();
        for (int i0 = 0; i0 < j; i0 += 32) {
            const int i = i0 + lane;
            if (i < j) out[(size_t)j * n + i] = row_distance<ELEM, METRIC>(q, g.vectors + (size_t)ids[i] * g.row_stride, nvec);
        }
    }
}

typedef void (*pp_kernel_t)(GraphView, const int32_t *, int, float *);
static pp_kernel_t pick_pp(int elem, int metric) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return pending_pairs_kernel<ELEM_F32, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return pending_pairs_kernel<ELEM_F32, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return pending_pairs_kernel<ELEM_F32, METRIC_L1>;
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return pending_pairs_kernel<ELEM_F16, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return pending_pairs_kernel<ELEM_F16, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return pending_pairs_kernel<ELEM_F16, METRIC_L1>;
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return pending_pairs_kernel<ELEM_BIT, METRIC_HAMMING>;
        if (metric == METRIC_JACCARD) return pending_pairs_kernel<ELEM_BIT, METRIC_JACCARD>;
    }
    return nullptr;
}

cudaError_t launch_pending_pairs(int elem, int metric, const GraphView &g, const int32_t *ids, int n, float *out, int sm_count,
                                 cudaStream_t stream) {
    pp_kernel_t k = pick_pp(elem, metric);
    if (!k) return cudaErrorInvalidValue;
    if (n <= 1) return cudaSuccess;
    const size_t smem = (size_t)(SEARCH_THREADS / 32) * g.row_stride;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = std::max(1, std::min((n + 3) / 4, sm_count * 4));
    k<<<grid, SEARCH_THREADS, smem, stream>>>(g, ids, n, out);
    return cudaGetLastError();
}

typedef void (*pair_kernel_t)(GraphView, PairParams);
static pair_kernel_t pick_pair(int elem, int metric) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return pairwise_kernel<ELEM_F32, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return pairwise_kernel<ELEM_F32, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return pairwise_kernel<ELEM_F32, METRIC_L1>;
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return pairwise_kernel<ELEM_F16, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return pairwise_kernel<ELEM_F16, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return pairwise_kernel<ELEM_F16, METRIC_L1>;
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return pairwise_kernel<ELEM_BIT, METRIC_HAMMING>;
        if (metric == METRIC_JACCARD) return pairwise_kernel<ELEM_BIT, METRIC_JACCARD>;
    }
    return nullptr;
}

cudaError_t launch_pairwise(int elem, int metric, const GraphView &g, const PairParams &p, int max_rows, int grid, cudaStream_t stream) {
    pair_kernel_t k = pick_pair(elem, metric);
    if (!k) return cudaErrorInvalidValue;
    if (p.n_groups <= 0) return cudaSuccess;
    const size_t smem = 128 + (size_t)max_rows * p.stage_stride;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, PAIR_THREADS, smem, stream>>>(g, p);
    return cudaGetLastError();
}

typedef void (*build_kernel_t)(GraphView, BuildParams);

template <int ELEM, int METRIC>
static build_kernel_t pick_list(int list_j, bool flagged) {
    if (flagged) return list_j == 0 ? hnsw_build_search_kernel<ELEM, METRIC, 0, true> : nullptr;
    switch (list_j) {
        case 2: return hnsw_build_search_kernel<ELEM, METRIC, 2>;
        case 4: return hnsw_build_search_kernel<ELEM, METRIC, 4>;
        case 8: return hnsw_build_search_kernel<ELEM, METRIC, 8>;
        case 0: return hnsw_build_search_kernel<ELEM, METRIC, 0>;
    }
    return nullptr;
}

static build_kernel_t pick_kernel(int elem, int metric, int list_j, bool flagged) {
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

cudaError_t build_kernel_occupancy(int elem, int metric, int list_j, bool flagged, size_t smem_block, int threads, int *blocks_per_sm) {
    build_kernel_t k = pick_kernel(elem, metric, list_j, flagged);
    if (!k) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_block);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads, smem_block);
}

cudaError_t launch_build_search(int elem, int metric, int list_j, const GraphView &g, const BuildParams &p, int grid,
                                size_t smem_block, int threads, cudaStream_t stream) {
    build_kernel_t k = pick_kernel(elem, metric, list_j, g.flags != nullptr);
    if (!k) return cudaErrorInvalidValue;
    k<<<grid, threads, smem_block, stream>>>(g, p);
    return cudaGetLastError();
}

typedef void (*repair_kernel_t)(GraphView, RepairParams);
static repair_kernel_t pick_repair(int elem, int metric) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return hnsw_repair_search_kernel<ELEM_F32, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return hnsw_repair_search_kernel<ELEM_F32, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return hnsw_repair_search_kernel<ELEM_F32, METRIC_L1>;
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return hnsw_repair_search_kernel<ELEM_F16, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return hnsw_repair_search_kernel<ELEM_F16, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return hnsw_repair_search_kernel<ELEM_F16, METRIC_L1>;
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return hnsw_repair_search_kernel<ELEM_BIT, METRIC_HAMMING>;
        if (metric == METRIC_JACCARD) return hnsw_repair_search_kernel<ELEM_BIT, METRIC_JACCARD>;
    }
    return nullptr;
}

cudaError_t repair_kernel_occupancy(int elem, int metric, size_t smem_block, int threads, int *blocks_per_sm) {
    repair_kernel_t k = pick_repair(elem, metric);
    if (!k) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_block);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads, smem_block);
}

cudaError_t launch_repair_search(int elem, int metric, const GraphView &g, const RepairParams &p, int grid, size_t smem_block,
                                 int threads, cudaStream_t stream) {
    repair_kernel_t k = pick_repair(elem, metric);
    if (!k) return cudaErrorInvalidValue;
    k<<<grid, threads, smem_block, stream>>>(g, p);
    return cudaGetLastError();
}

cudaError_t launch_scatter_rows(int32_t *dst, const int64_t *row_offsets, const int32_t *packed, int row_len, int n_rows,
                                cudaStream_t stream) {
    if (n_rows <= 0) return cudaSuccess;
    int total = n_rows * row_len;
    scatter_rows_kernel<<<(total + 255) / 256, 256, 0, stream>>>(dst, row_offsets, packed, row_len, n_rows);
    return cudaGetLastError();
}

}  // namespace pgvrx
