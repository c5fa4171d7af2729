// hnsw_team.cu -- K2t: the build-side search (find_element_neighbors, src/graph/mod.rs:355-427) with one CTA per
// pending insert instead of one warp.
//
// A build round is a few dozen searches against one snapshot, and the round cannot commit before the slowest of them
// has finished: the build is bound by the latency of ONE search, and a search is a chain of expansions
// each of which waits for two dependent memory round trips (the candidate's neighbour row, then the neighbours' vector
// rows) and a sequential distance chain.  The chain itself cannot be cut -- which candidate is expanded next depends on
// the previous expansion -- but what an expansion NEEDS from memory does not depend on the search state at all: the
// distances from the query to the entries of a candidate's neighbour list are a pure function of (query, candidate,
// layer) on a frozen snapshot.  So the CTA splits the work:
//
//   * helper warps compute "expansions" ahead of time for the nearest not-yet-expanded candidates of the result list:
//     neighbour row -> TMA row gather -> one distance per lane (the same sequential arithmetic as everywhere else) ->
//     (ids, distances, levels) into a cache slot in shared memory;
//   * the master warp runs search_layer's sequential semantics exactly as hnsw_device.cuh does -- pop the nearest
//     unexpanded candidate, visited test-and-set in list order BEFORE the level test (graph/mod.rs:206-216), accept /
//     evict in list order (graph/mod.rs:226-242), expansion log for the host's replay rule -- but takes the distances
//     from the cache, and keeps the visited set in shared memory.  On a miss it computes the expansion itself.
//
// Speculation is almost free of waste: nearly every member of the final W gets expanded (n_expand ~ 1.2 ef), so an
// expansion computed for a candidate that currently sits in W is used unless the candidate is evicted first.  Results
// are identical to the one-warp kernel by construction (same arithmetic, same order of state changes); a search that
// exhausts the shared-memory visited table or the tie tail reports a status and the host re-runs it on the one-warp
// kernel with large scratch (build_host.cu's escalation levels).
#include <algorithm>

#include "hnsw_device.cuh"
#include "hnsw_kernels.h"

namespace pgvrx {

__host__ __device__ inline uint32_t round16t(uint32_t x) { return (x + 15u) & ~15u; }

constexpr int TEAM_QUEUE = 64;          // request ring (>= cache slots + 1)
constexpr int TEAM_WANT = 6;            // candidates requested ahead of the master

struct TeamCtl {
    int q_head;                         // tickets handed to helpers (atomicAdd)
    int q_tail;                         // requests published by the master
    int quit;                           // the CTA has no more searches
    int layer;                          // layer of the current requests
    int q_slot[TEAM_QUEUE];
    int slot_id[TEAM_MAX_SLOTS + 1];    // candidate id, -1 = none (slot TEAM_MAX_SLOTS is the master's own)
    int slot_state[TEAM_MAX_SLOTS + 1]; // 0 free, 1 requested / in progress, 2 ready
    int slot_cnt[TEAM_MAX_SLOTS + 1];   // valid entries
};

__device__ __forceinline__ int ld_volatile(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_volatile(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }

size_t team_smem_bytes(const TeamGeom &t) {
    size_t b = (sizeof(TeamCtl) + 127) & ~(size_t)127;
    b += t.row_stride;                                                   // query
    b += (size_t)t.vis_entries * 4;                                      // visited table
    b += (size_t)(t.n_slots + 1) * t.lmp * (4 + 4);                      // slot ids + distances
    b += round16t((uint32_t)((t.n_slots + 1) * t.lmp));                  // slot levels
    if (t.list_j == 0) b += round16t((uint32_t)((t.ef + t.tail_cap + 1) * 8));
    b += (size_t)t.warps * ((size_t)t.stage_bytes + 128 + 16);           // per warp: staging, neighbour ids, mbarrier
    return b;
}

// one expansion: distances from the query to every entry of (cid, layer)'s neighbour list, in list order
template <int ELEM, int METRIC>
__device__ __forceinline__ void team_expand(WarpCtx &c, const GraphView &g, int layer, uint32_t cid, int32_t *o_nbr, float *o_dist,
                                            uint8_t *o_lvl, int *o_cnt) {
    const int lane = c.lane;
    const int lm = layer == 0 ? 2 * g.m : g.m;
    const int32_t *row = layer == 0 ? g.adj0 + (size_t)cid * (size_t)lm
                                    : g.adj_up + ((size_t)__ldg(g.upper_base + cid) + (size_t)(layer - 1)) * (size_t)lm;
    int total = 0;
    for (int base = 0; base < lm; base += 32) {
        const int j = base + lane;
        const int32_t nb = j < lm ? __ldg(row + j) : -1;
        const bool valid = nb >= 0;                       // lists are dense prefixes (scan.rs:273-279)
        const uint32_t vm = __ballot_sync(FULL, valid);
        if (!vm) break;
        const int cnt = __popc(vm);
        int lvl = 255;
        if (layer > 0 && valid) lvl = __ldg(g.level + nb);
        if (valid) c.nbr[lane] = nb;
        __syncwarp();
        const float d = eval_neighbours<ELEM, METRIC>(c, g, cnt, valid, false);
        if (valid) { o_nbr[j] = nb; o_dist[j] = d; o_lvl[j] = (uint8_t)lvl; }
        total += cnt;
        __syncwarp();                                     // c.nbr is rewritten by the next chunk
    }
    if (lane == 0) *o_cnt = total;
}

// first `kmax` unexpanded candidates of the list in list order (ids, uniform across the warp); returns how many
template <int J>
__device__ __forceinline__ int peek_unexpanded(const RegList<J> &S, const WarpCtx &c, uint32_t *out, int kmax) {
    uint32_t b[J], any = 0;
#pragma unroll
    for (int j = 0; j < J; j++) { b[j] = __ballot_sync(FULL, S.k[j] != KEY_NONE && (S.k[j] & 1ull) == 0); any |= b[j]; }
    int n = 0;
    while (any && n < kmax) {
        const int l = __ffs(any) - 1;
        any &= any - 1;
#pragma unroll
        for (int j = 0; j < J; j++) {
            if (((b[j] >> l) & 1u) && n < kmax) {
                const uint64_t key = __shfl_sync(FULL, S.k[j], l);
                out[n++] = key_id(key);
            }
        }
    }
    return n;
}
__device__ __forceinline__ int peek_unexpanded(const SmemList &S, const WarpCtx &c, uint32_t *out, int kmax) {
    int n = 0;
    for (int base = (S.hint >> 5) << 5; base < c.len && n < kmax; base += 32) {
        const int i = base + c.lane;
        const bool un = i >= S.hint && i < c.len && (S.S[i] & 1ull) == 0;
        uint32_t b = __ballot_sync(FULL, un);
        while (b && n < kmax) {
            const int l = __ffs(b) - 1;
            b &= b - 1;
            out[n++] = key_id(S.S[base + l]);
        }
    }
    return n;
}

struct TeamMaster {
    TeamCtl *ctl;
    uint32_t *vis;                 // shared-memory visited table
    int32_t *s_nbr;
    float *s_dist;
    uint8_t *s_lvl;
    int n_slots, lmp;
};

// wait until no request is in flight, then drop every cached expansion (new search / new layer)
__device__ __forceinline__ void team_flush(const TeamMaster &T, const WarpCtx &c, int layer) {
    const int lane = c.lane;
    if (lane < T.n_slots) {
        while (ld_volatile(&T.ctl->slot_state[lane]) == 1) __nanosleep(32);
        T.ctl->slot_state[lane] = 0;
        T.ctl->slot_id[lane] = -1;
    }
    if (lane == 0) T.ctl->layer = layer;
    __threadfence_block();
    __syncwarp();
}

// ask the helpers for the expansions of the nearest unexpanded candidates that are not cached yet
template <class LIST>
__device__ __forceinline__ void team_request(const TeamMaster &T, const LIST &S, const WarpCtx &c) {
    uint32_t want[TEAM_WANT];
    const int n = peek_unexpanded(S, c, want, TEAM_WANT);
    const int lane = c.lane;
    for (int t = 0; t < n; t++) {
        const int sid = lane < T.n_slots ? T.ctl->slot_id[lane] : -1;
        const int sst = lane < T.n_slots ? ld_volatile(&T.ctl->slot_state[lane]) : 1;
        if (__ballot_sync(FULL, sst != 0 && sid == (int)want[t])) continue;            // cached or on its way
        uint32_t freeb = __ballot_sync(FULL, sst == 0);
        if (!freeb) {
            // every slot is taken: drop finished expansions of candidates that are no longer among the wanted ones
            // (evicted from the list, or simply further down than TEAM_WANT); in-flight ones cannot be dropped
            bool keep = false;
            for (int u = 0; u < n; u++) keep = keep || sid == (int)want[u];
            const bool drop = lane < T.n_slots && sst == 2 && !keep;
            if (drop) { T.ctl->slot_state[lane] = 0; T.ctl->slot_id[lane] = -1; }
            freeb = __ballot_sync(FULL, drop);
            if (!freeb) return;
        }
        const int s = __ffs(freeb) - 1;
        if (lane == 0) {
            T.ctl->slot_id[s] = (int)want[t];
            T.ctl->slot_state[s] = 1;
            const int tail = T.ctl->q_tail;
            T.ctl->q_slot[tail % TEAM_QUEUE] = s;
            __threadfence_block();
            st_volatile(&T.ctl->q_tail, tail + 1);
        }
        __syncwarp();
    }
}

// search_layer (graph/mod.rs:161-255) by the master warp; see hnsw_device.cuh search_layer for the rules, which are the
// same line for line -- only the source of the distances and the home of the visited set differ
template <int ELEM, int METRIC, class LIST>
__device__ __forceinline__ void team_search_layer(const TeamMaster &T, WarpCtx &c, LIST &S, const GraphView &g, int layer,
                                                  uint32_t vis_mask, uint32_t vis_limit) {
    const int lane = c.lane;
    uint32_t vis_count = (uint32_t)c.len;
    team_request(T, S, c);
    while (true) {
        uint64_t ckey;
        if (!S.pop_nearest(c, ckey)) break;   // every remaining candidate would fail `c.distance > f_dist`
        const uint32_t cid = key_id(ckey);
        if (lane == 0 && c.rs_n < c.rs_cap) {
            int32_t *e = c.rs_log + (size_t)c.rs_ent * (size_t)c.rs_n;
            e[0] = (int32_t)cid;
            e[1] = c.len >= c.ef ? (int32_t)c.f_ord : -1;
            e[2] = layer;
        }
        c.rs_n++;
        const bool log_full = c.len >= c.ef;        // W full at the start of this expansion
        const uint32_t log_f = c.f_ord;
        if (layer == 0) c.n_exp0++; else c.n_expu++;
        // the candidate's expansion: cached, on its way, or computed here
        int s;
        {
            const int sid = lane < T.n_slots ? T.ctl->slot_id[lane] : -1;
            const int sst = lane < T.n_slots ? ld_volatile(&T.ctl->slot_state[lane]) : 0;
            const uint32_t hit = __ballot_sync(FULL, sst != 0 && sid == (int)cid);
            if (hit) {
                s = __ffs(hit) - 1;
                while (ld_volatile(&T.ctl->slot_state[s]) != 2) __nanosleep(20);
            } else {
                s = TEAM_MAX_SLOTS;
                team_expand<ELEM, METRIC>(c, g, layer, cid, T.s_nbr + (size_t)T.n_slots * T.lmp, T.s_dist + (size_t)T.n_slots * T.lmp,
                                          T.s_lvl + (size_t)T.n_slots * T.lmp, &T.ctl->slot_cnt[TEAM_MAX_SLOTS]);
            }
            __threadfence_block();
            __syncwarp();
        }
        const int srow = s == TEAM_MAX_SLOTS ? T.n_slots : s;
        const int total = ld_volatile(&T.ctl->slot_cnt[s]);
        const int32_t *e_nbr = T.s_nbr + (size_t)srow * T.lmp;
        const float *e_dist = T.s_dist + (size_t)srow * T.lmp;
        const uint8_t *e_lvl = T.s_lvl + (size_t)srow * T.lmp;
        for (int base = 0; base < total; base += 32) {
            const int j = base + lane;
            const bool valid = j < total;
            const int32_t nb = valid ? e_nbr[j] : -1;
            const float d = valid ? e_dist[j] : 0.0f;
            const int lvl = valid ? (int)e_lvl[j] : 255;
            // the host's replay rule needs: which entries of this list are nearer than W's furthest distance was when
            // the expansion started (every valid entry has its distance, visited or not)
            if (c.rs_n <= c.rs_cap) {
                const uint32_t nm = __ballot_sync(FULL, valid && log_full && f32_ord(d) < log_f);
                if (lane == 0) c.rs_log[(size_t)c.rs_ent * (size_t)(c.rs_n - 1) + 3 + (base >> 5)] = (int32_t)nm;
            }
            // visited BEFORE the level test (graph/mod.rs:206-216)
            const bool fresh = valid && vis_insert(T.vis, vis_mask, (uint32_t)nb);
            const uint32_t fm = __ballot_sync(FULL, fresh);
            vis_count += __popc(fm);
            c.n_vis += __popc(fm);
            if (vis_count > vis_limit) { c.status |= ST_VIS_OVERFLOW; break; }
            const bool pass = fresh && lvl >= layer;                                  // graph/mod.rs:214
            c.n_dist += __popc(__ballot_sync(FULL, pass));
            // accept in neighbour order (graph/mod.rs:226-242)
            const uint32_t dord = f32_ord(d);
            uint32_t am = __ballot_sync(FULL, pass && (c.len < c.ef || dord < c.f_ord));
            while (am) {
                const int r = __ffs(am) - 1;
                am &= am - 1;
                const float dr = __shfl_sync(FULL, d, r);
                const int32_t idr = __shfl_sync(FULL, nb, r);
                if (c.len < c.ef || f32_ord(dr) < c.f_ord) S.insert(c, make_key(dr, (uint32_t)idr));
            }
        }
        // (log words of 32-entry chunks beyond the list's length are never written, and never read: the host only looks
        // at positions below the list length it saw at the snapshot)
        if (s != TEAM_MAX_SLOTS && lane == 0) { T.ctl->slot_id[s] = -1; st_volatile(&T.ctl->slot_state[s], 0); }
        __syncwarp();
        if (c.status) return;
        team_request(T, S, c);
    }
}

template <int J> struct TListOf {
    typedef RegList<J> type;
    __device__ static __forceinline__ void attach(type &, uint8_t *) {}
};
template <> struct TListOf<0> {
    typedef SmemList type;
    __device__ static __forceinline__ void attach(type &l, uint8_t *mem) { l.S = reinterpret_cast<uint64_t *>(mem); l.hint = 0; }
};

template <int ELEM, int METRIC, int LISTJ>
__global__ void __launch_bounds__(TEAM_MAX_WARPS * 32) hnsw_team_build_kernel(GraphView g, BuildParams p, TeamGeom t) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // ---- carve shared memory (team_smem_bytes is the same walk) ----
    uint8_t *ptr = smem;
    TeamCtl *ctl = reinterpret_cast<TeamCtl *>(ptr);
    ptr += (sizeof(TeamCtl) + 127) & ~(size_t)127;
    uint8_t *qsm = ptr;
    ptr += g.row_stride;
    uint32_t *vis = reinterpret_cast<uint32_t *>(ptr);
    ptr += (size_t)t.vis_entries * 4;
    int32_t *s_nbr = reinterpret_cast<int32_t *>(ptr);
    ptr += (size_t)(t.n_slots + 1) * t.lmp * 4;
    float *s_dist = reinterpret_cast<float *>(ptr);
    ptr += (size_t)(t.n_slots + 1) * t.lmp * 4;
    uint8_t *s_lvl = ptr;
    ptr += round16t((uint32_t)((t.n_slots + 1) * t.lmp));
    uint8_t *s_mem = ptr;
    const int cap = p.efc + p.tail_cap;
    if (LISTJ == 0) ptr += round16t((uint32_t)((cap + 1) * 8));
    uint8_t *wbase = ptr + (size_t)warp * ((size_t)t.stage_bytes + 128 + 16);

    WarpCtx c;
    c.stage = wbase;
    c.qsm = qsm;
    c.nbr = reinterpret_cast<int32_t *>(wbase + t.stage_bytes);
    c.mbar = reinterpret_cast<uint64_t *>(wbase + t.stage_bytes + 128);
    c.stage_stride = p.stage_stride;
    c.R = p.R;
    c.chunk_bytes = p.chunk_bytes;
    c.chunk_stride = p.chunk_stride;
    c.nvec = (int)(g.row_stride / 16);
    c.cap = cap;
    c.lane = lane;
    c.phase = 0;
    c.phase2 = 0;
    c.stage2 = (c.chunk_bytes != 0 && (uint32_t)t.stage_bytes >= 64u * c.chunk_stride) ? c.stage + 32u * (size_t)c.chunk_stride : nullptr;
    c.policy = l2_policy_evict_first();
    c.vis = nullptr;
    c.spec = true;
    c.hybrid = false;
    if (lane == 0) {
        mbar_init(c.mbar, 1);
        mbar_init(c.mbar + 1, 1);
        fence_mbar_init();
    }
    if (threadIdx.x == 0) {
        ctl->q_head = ctl->q_tail = ctl->quit = 0;
        ctl->layer = 0;
    }
    if (threadIdx.x <= TEAM_MAX_SLOTS) { ctl->slot_id[threadIdx.x] = -1; ctl->slot_state[threadIdx.x] = 0; ctl->slot_cnt[threadIdx.x] = 0; }
    __syncthreads();

    if (warp != 0) {
        // ---------------- helpers: serve expansion requests until the master quits ----------------
        while (true) {
            int ticket = 0;
            if (lane == 0) ticket = atomicAdd(&ctl->q_head, 1);
            ticket = __shfl_sync(FULL, ticket, 0);
            bool quit = false;
            while (ld_volatile(&ctl->q_tail) <= ticket) {
                if (ld_volatile(&ctl->quit)) { quit = true; break; }
                __nanosleep(40);
            }
            if (quit) break;
            __threadfence_block();
            const int s = ld_volatile(&ctl->q_slot[ticket % TEAM_QUEUE]);
            const uint32_t cid = (uint32_t)ld_volatile(&ctl->slot_id[s]);
            const int layer = ld_volatile(&ctl->layer);
            team_expand<ELEM, METRIC>(c, g, layer, cid, s_nbr + (size_t)s * t.lmp, s_dist + (size_t)s * t.lmp, s_lvl + (size_t)s * t.lmp,
                                      &ctl->slot_cnt[s]);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) st_volatile(&ctl->slot_state[s], 2);
        }
        return;
    }

    // ---------------- master ----------------
    TeamMaster T;
    T.ctl = ctl; T.vis = vis; T.s_nbr = s_nbr; T.s_dist = s_dist; T.s_lvl = s_lvl; T.n_slots = t.n_slots; T.lmp = t.lmp;
    typename TListOf<LISTJ>::type S;
    TListOf<LISTJ>::attach(S, s_mem);
    const uint32_t up_entries = t.vis_entries < 1024u ? t.vis_entries : 1024u;

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
        c.rs_cap = p.rs_cap;
        c.rs_n = 0;
        for (int l = lane; l < p.max_layers; l += 32) p.w_counts[(size_t)w * p.max_layers + l] = -1;

        team_flush(T, c, 0);                              // helpers idle: the query row may change
        {
            const uint4 *q4 = reinterpret_cast<const uint4 *>(g.vectors + (size_t)e * g.row_stride);
            uint4 *d4 = reinterpret_cast<uint4 *>(c.qsm);
            for (int i = lane; i < c.nvec; i += 32) d4[i] = __ldg(q4 + i);
        }
        if (lane == 0) c.nbr[0] = g.entry;
        __threadfence_block();
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
            team_flush(T, c, lc);
            vis_clear(vis, up_entries, lane);
            if (lane == 0) vis_insert(vis, up_entries - 1, key_id(ep_key));
            __syncwarp();
            c.n_vis++;
            S.reset1(c, ep_key);
            team_search_layer<ELEM, METRIC>(T, c, S, g, lc, up_entries - 1, up_entries / 4 * 3 - 32);
            ep_key = S.get(c, 0) & ~1ull;                                           // :396-398 w[0]
        }
        // phase 2: ef_construction search on every layer the new element joins (graph/mod.rs:403-426)
        c.ef = p.efc;
        bool first = true;
        for (int lc = min(lvl, g.entry_level); lc >= 0 && c.status == 0; lc--) {
            team_flush(T, c, lc);
            vis_clear(vis, t.vis_entries, lane);
            if (first) {
                if (lane == 0) vis_insert(vis, t.vis_entries - 1, key_id(ep_key));
                __syncwarp();
                c.n_vis++;
                S.reset1(c, ep_key);
            } else {
                S.reopen(c, vis, t.vis_entries - 1);                                // :425 ep = w
                c.n_vis += c.len;
            }
            first = false;
            team_search_layer<ELEM, METRIC>(T, c, S, g, lc, t.vis_entries - 1, t.vis_entries / 4 * 3 - 32);
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
    team_flush(T, c, 0);
    if (lane == 0) { __threadfence_block(); st_volatile(&ctl->quit, 1); }
}

typedef void (*team_kernel_t)(GraphView, BuildParams, TeamGeom);

template <int ELEM, int METRIC>
static team_kernel_t pick_team_list(int list_j) {
    switch (list_j) {
        case 2: return hnsw_team_build_kernel<ELEM, METRIC, 2>;
        case 4: return hnsw_team_build_kernel<ELEM, METRIC, 4>;
        case 8: return hnsw_team_build_kernel<ELEM, METRIC, 8>;
        case 0: return hnsw_team_build_kernel<ELEM, METRIC, 0>;
    }
    return nullptr;
}

static team_kernel_t pick_team(int elem, int metric, int list_j) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return pick_team_list<ELEM_F32, METRIC_L2SQ>(list_j);
        if (metric == METRIC_NEG_IP) return pick_team_list<ELEM_F32, METRIC_NEG_IP>(list_j);
        if (metric == METRIC_L1) return pick_team_list<ELEM_F32, METRIC_L1>(list_j);
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return pick_team_list<ELEM_F16, METRIC_L2SQ>(list_j);
        if (metric == METRIC_NEG_IP) return pick_team_list<ELEM_F16, METRIC_NEG_IP>(list_j);
        if (metric == METRIC_L1) return pick_team_list<ELEM_F16, METRIC_L1>(list_j);
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return pick_team_list<ELEM_BIT, METRIC_HAMMING>(list_j);
        if (metric == METRIC_JACCARD) return pick_team_list<ELEM_BIT, METRIC_JACCARD>(list_j);
    }
    return nullptr;
}

cudaError_t team_kernel_occupancy(int elem, int metric, int list_j, size_t smem_block, int threads, int *blocks_per_sm) {
    team_kernel_t k = pick_team(elem, metric, list_j);
    if (!k) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_block);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads, smem_block);
}

cudaError_t launch_team_build(int elem, int metric, int list_j, const GraphView &g, const BuildParams &p, const TeamGeom &t, int grid,
                              size_t smem_block, cudaStream_t stream) {
    team_kernel_t k = pick_team(elem, metric, list_j);
    if (!k) return cudaErrorInvalidValue;
    k<<<grid, t.warps * 32, smem_block, stream>>>(g, p, t);
    return cudaGetLastError();
}

}  // namespace pgvrx
