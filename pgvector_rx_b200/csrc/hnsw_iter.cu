// hnsw_iter.cu -- K3: iterative index scan (hnsw.iterative_scan = relaxed_order / strict_order):
// get_scan_items on the first round, resume_scan_items on later rounds (src/index/scan.rs:458-577), and
// amgettuple's batch-level emission rules (scan.rs:794-875) for a batch of queries, one warp per query.
//
// What persists per query between rounds lives in HBM:
//   visited   -- a bitmap over element ids (scan.rs:601-604 keeps one HashSet per scan and never clears it);
//   discarded -- the reference's unbounded min-heap (scan.rs:340-345, 385-405, 421-438) as unordered 64-key
//                chunks (each with its minimum) in extents drawn from a shared pool; only its contents as a
//                multiset matter, because it is read solely by "pop the ef_search nearest" at the next resume
//                (scan.rs:552-561);
//   tuples, previous_distance (scan.rs:801-806, 871).
// The reference pushes an element to `discarded` when it is evicted from W and -- if it was never expanded --
// once more when it is popped at the stop test or drained from C at the end (it stays in C after the
// eviction); both pushes are reproduced, so resumed rounds see the same duplicates.
#include "hnsw_device.cuh"
#include "hnsw_kernels.h"

namespace pgvrx {

// ---- discarded: chunks with a per-chunk minimum, allocated in extents ------------------------------------
// A query's keys live in 64-key chunks; chunks are handed out in EXTENTS of 32 consecutive chunk ids from a shared
// pool, and a query's extents form a chain.  Every chunk carries its fill count and the minimum of its keys, so
// "pop the ef nearest" reads one 256-byte line of minima per extent (2048 keys) and opens only the chunks whose
// minimum is below the running ef-th smallest key: O(n / 2048 + ef ln(n / ef)) warp steps instead of two passes over
// all n keys -- `discarded` grows by ~n_dist keys per round and a filtered scan runs hundreds of rounds.
struct DiscWriter {
    uint64_t *pool;      // n_ext x ITER_EXT x ITER_CHUNK keys
    uint64_t *cmin;      // n_ext x ITER_EXT: smallest key of the chunk (KEY_NONE when empty)
    int *cfill;          // n_ext x ITER_EXT: keys in the chunk
    int *next;           // n_ext: next extent of the chain or -1
    int *pool_next;      // bump allocator (extents)
    int pool_cap;        // extents
    int head;            // first extent or -1
    int tail;            // chunk that takes appends or -1
    int fill, live;      // keys in the tail chunk, live keys in the chain
    uint64_t tail_min;   // running minimum of the tail chunk (written back when the tail moves on / at the end of the round)
    uint32_t status;
};

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_min_sync(FULL, hi);
    const uint32_t ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xFFFFFFFFu);
    return ((uint64_t)mh << 32) | ml;
}

// a fresh extent: every chunk empty
__device__ __forceinline__ int disc_alloc_ext(DiscWriter &d, int lane) {
    int e = -1;
    if (lane == 0) {
        e = atomicAdd(d.pool_next, 1);
        if (e >= d.pool_cap) e = -1; else d.next[e] = -1;
    }
    e = __shfl_sync(FULL, e, 0);
    if (e < 0) { d.status |= ST_POOL_OVERFLOW; return -1; }
    d.cmin[(size_t)e * ITER_EXT + lane] = KEY_NONE;
    d.cfill[(size_t)e * ITER_EXT + lane] = 0;
    return e;
}

// chunk that follows the tail chunk: the next one of its extent, or the first of a fresh extent linked behind it
__device__ __forceinline__ int disc_next_chunk(DiscWriter &d, int lane) {
    if ((d.tail + 1) % ITER_EXT != 0) return d.tail + 1;
    const int e = disc_alloc_ext(d, lane);
    if (e < 0) return -1;
    if (lane == 0) d.next[d.tail / ITER_EXT] = e;
    return e * ITER_EXT;
}

// the tail chunk's metadata -> memory (the chunk stays the tail)
__device__ __forceinline__ void disc_flush_tail(const DiscWriter &d, int lane) {
    if (d.tail >= 0 && lane == 0) { d.cmin[d.tail] = d.tail_min; d.cfill[d.tail] = d.fill; }
}

// warp-collective append of the keys of the lanes with pred set
__device__ __forceinline__ void disc_append(DiscWriter &d, bool pred, uint64_t key, int lane) {
    const uint32_t m = __ballot_sync(FULL, pred);
    if (!m || (d.status & ST_POOL_OVERFLOW)) return;
    const int n = __popc(m);
    const int rank = __popc(m & ((1u << lane) - 1u));
    if (d.tail < 0) {
        const int e = disc_alloc_ext(d, lane);
        if (e < 0) return;
        d.head = e;
        d.tail = e * ITER_EXT;
        d.fill = 0;
        d.tail_min = KEY_NONE;
    }
    const bool spill = d.fill + n > ITER_CHUNK;
    int nxt = -1;
    if (spill) {
        nxt = disc_next_chunk(d, lane);
        if (nxt < 0) return;
    }
    const int pos = d.fill + rank;
    const bool first = pred && pos < ITER_CHUNK;
    if (pred) {
        if (first) d.pool[(size_t)d.tail * ITER_CHUNK + pos] = key;
        else d.pool[(size_t)nxt * ITER_CHUNK + (pos - ITER_CHUNK)] = key;
    }
    const uint64_t m1 = warp_min_u64(first ? key : KEY_NONE);
    if (m1 < d.tail_min) d.tail_min = m1;
    if (spill) {
        if (lane == 0) { d.cmin[d.tail] = d.tail_min; d.cfill[d.tail] = ITER_CHUNK; }     // the old tail is full
        d.tail_min = warp_min_u64(pred && !first ? key : KEY_NONE);
        d.tail = nxt;
        d.fill = d.fill + n - ITER_CHUNK;
    } else {
        d.fill += n;
    }
    d.live += n;
    __syncwarp();
}

// warp-cooperative insert into an ascending list of capacity k, keeping the k smallest (multiset semantics)
__device__ __forceinline__ void sel_insert(uint64_t *L, int &len, int k, uint64_t key, int lane) {
    if (len == k && key >= L[k - 1]) return;
    int pos = 0;
    for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        const bool lt = i < len && L[i] < key;
        pos += __popc(__ballot_sync(FULL, lt));
    }
    const int newlen = len < k ? len + 1 : k;
    for (int base = ((newlen - 1) >> 5) << 5; base >= 0 && base + 32 > pos; base -= 32) {
        const int i = base + lane;
        const bool mv = i >= pos && i < newlen - 1;
        const uint64_t v = mv ? L[i] : 0;
        __syncwarp();
        if (mv) L[i + 1] = v;
        __syncwarp();
    }
    if (lane == 0) L[pos] = key;
    __syncwarp();
    len = newlen;
}

// discarded.pop() x ef_search (scan.rs:552-561): the `ef` smallest keys go to L[0, n_sel) ascending and are removed
// from their chunks (each compacted in place, minimum and fill count updated).  Returns n_sel.  Called at the start of
// a round, before anything is appended: the tail chunk's metadata in memory is current.
__device__ __forceinline__ int disc_pop_nearest(DiscWriter &d, uint64_t *L, int ef, int lane) {
    if (d.live == 0 || d.head < 0) return 0;
    int n_sel = 0;
    // pass 1: select.  A chunk is opened only when its minimum beats the current ef-th smallest key.
    for (int ext = d.head; ext >= 0;) {
        int nx = 0;
        if (lane == 0) nx = d.next[ext];
        const uint64_t mn = d.cmin[(size_t)ext * ITER_EXT + lane];
        const int fl = d.cfill[(size_t)ext * ITER_EXT + lane];
        nx = __shfl_sync(FULL, nx, 0);
        uint32_t cm = __ballot_sync(FULL, mn < (n_sel == ef ? L[ef - 1] : KEY_NONE));
        while (cm) {
            const int r = __ffs(cm) - 1;
            cm &= cm - 1;
            const uint64_t mnr = __shfl_sync(FULL, mn, r);
            const int flr = __shfl_sync(FULL, fl, r);
            if (!(mnr < (n_sel == ef ? L[ef - 1] : KEY_NONE))) continue;      // the threshold has tightened meanwhile
            const uint64_t *ck = d.pool + ((size_t)ext * ITER_EXT + r) * ITER_CHUNK;
            for (int base = 0; base < flr; base += 32) {
                const int i = base + lane;
                const uint64_t key = i < flr ? ck[i] : KEY_NONE;
                uint32_t am = __ballot_sync(FULL, key < (n_sel == ef ? L[ef - 1] : KEY_NONE));
                while (am) {
                    const int rr = __ffs(am) - 1;
                    am &= am - 1;
                    sel_insert(L, n_sel, ef, __shfl_sync(FULL, key, rr), lane);
                }
            }
        }
        ext = nx;
    }
    if (n_sel == 0) return 0;
    // pass 2: remove the selected multiset (everything below T, and as many copies of T as were selected) from the
    // chunks that can hold any of it
    const uint64_t T = L[n_sel - 1];
    int eq_left = 0;
    for (int base = 0; base < n_sel; base += 32) {
        const int i = base + lane;
        eq_left += __popc(__ballot_sync(FULL, i < n_sel && L[i] == T));
    }
    int to_remove = n_sel;
    for (int ext = d.head; ext >= 0 && to_remove > 0;) {
        int nx = 0;
        if (lane == 0) nx = d.next[ext];
        const uint64_t mn = d.cmin[(size_t)ext * ITER_EXT + lane];
        const int fl = d.cfill[(size_t)ext * ITER_EXT + lane];
        nx = __shfl_sync(FULL, nx, 0);
        uint32_t cm = __ballot_sync(FULL, mn <= T);
        while (cm && to_remove > 0) {
            const int r = __ffs(cm) - 1;
            cm &= cm - 1;
            const int flr = __shfl_sync(FULL, fl, r);
            const int chunk = ext * ITER_EXT + r;
            uint64_t *ck = d.pool + (size_t)chunk * ITER_CHUNK;
            int wfill = 0;
            uint64_t newmin = KEY_NONE;
            for (int base = 0; base < flr; base += 32) {
                const int i = base + lane;
                const bool valid = i < flr;
                const uint64_t key = valid ? ck[i] : KEY_NONE;
                const uint32_t eqm = __ballot_sync(FULL, valid && key == T);
                const int take = min(__popc(eqm), eq_left);
                const bool eq_removed = valid && key == T && __popc(eqm & ((1u << lane) - 1u)) < take;
                eq_left -= take;
                const bool keep = valid && !(key < T) && !eq_removed;
                const uint32_t km = __ballot_sync(FULL, keep);
                const int rank = __popc(km & ((1u << lane) - 1u));
                __syncwarp();
                if (keep) ck[wfill + rank] = key;              // compaction only moves keys towards the front
                const uint64_t mk = warp_min_u64(keep ? key : KEY_NONE);
                if (mk < newmin) newmin = mk;
                wfill += __popc(km);
                __syncwarp();
            }
            to_remove -= flr - wfill;
            if (lane == 0) { d.cmin[chunk] = newmin; d.cfill[chunk] = wfill; }
            if (chunk == d.tail) { d.fill = wfill; d.tail_min = newmin; }
        }
        ext = nx;
    }
    d.live -= n_sel;
    __syncwarp();
    return n_sel;
}

// ---- S in shared memory with the iterative scan's discard hooks -----------------------------------------
struct IterList {
    uint64_t *S;
    int hint;

    __device__ __forceinline__ uint64_t get(const WarpCtx &c, int i) const { return S[i]; }
    __device__ __forceinline__ bool pop_nearest(WarpCtx &c, uint64_t &key) {
        for (int base = (hint >> 5) << 5; base < c.len; base += 32) {
            const int i = base + c.lane;
            const bool un = i >= hint && i < c.len && (S[i] & 1ull) == 0;
            const uint32_t b = __ballot_sync(FULL, un);
            if (b) {
                const int idx = base + __ffs(b) - 1;
                key = S[idx];
                __syncwarp();
                if (c.lane == 0) S[idx] = key | 1ull;
                __syncwarp();
                hint = idx + 1;
                return true;
            }
        }
        return false;
    }
    __device__ __forceinline__ void insert(WarpCtx &c, uint64_t key, DiscWriter &d) {
        const int lane = c.lane;
        int pos = 0;
        for (int base = 0; base < c.len; base += 32) {
            const int i = base + lane;
            const bool lt = i < c.len && S[i] < key;
            pos += __popc(__ballot_sync(FULL, lt));
        }
        for (int base = ((c.len - 1) >> 5) << 5; base >= 0 && base + 31 >= pos && c.len > 0; base -= 32) {
            const int i = base + lane;
            const bool mv = i >= pos && i < c.len;
            const uint64_t v = mv ? S[i] : 0;
            __syncwarp();
            if (mv) S[i + 1] = v;
            __syncwarp();
        }
        if (lane == 0) S[pos] = key;
        __syncwarp();
        c.len += 1;
        if (pos < hint) hint = pos;
        if (c.len >= c.ef) {
            const uint32_t old_f = c.f_ord;
            const bool was_full = c.len > c.ef;
            c.f_ord = key_ord(S[c.ef - 1]);
            if (was_full) {
                // the element that just left W is pushed to discarded (scan.rs:421-429); it stays in C
                disc_append(d, lane == 0, S[c.ef] & ~1ull, lane);
                if (c.f_ord != old_f) {
                    // nothing behind W ties with W's furthest distance any more, so none of it can be expanded:
                    // the unexpanded part is what the stop test and the final drain of C push (scan.rs:340-345,434-438)
                    for (int base = c.ef; base < c.len; base += 32) {
                        const int i = base + lane;
                        const bool un = i < c.len && (S[i] & 1ull) == 0;
                        disc_append(d, un, un ? S[i] : 0, lane);
                    }
                    c.len = c.ef;
                } else if (c.len > c.cap) {
                    c.len = c.cap;
                    c.status |= ST_TIE_OVERFLOW;
                }
            }
        }
    }
};

// search_layer_disk at layer 0 with shared visited / discarded (scan.rs:302-448)
template <int ELEM, int METRIC>
__device__ __forceinline__ void search_layer_iter(WarpCtx &c, IterList &S, const GraphView &g, uint32_t *bitmap, DiscWriter &d) {
    const int lane = c.lane;
    const int lm = 2 * g.m;
    while (true) {
        uint64_t ckey;
        if (!S.pop_nearest(c, ckey)) break;
        const uint32_t cid = key_id(ckey);
        const int32_t *row = g.adj0 + (size_t)cid * (size_t)lm;
        c.n_exp0++;
        for (int base = 0; base < lm; base += 32) {
            const int j = base + lane;
            const int32_t nb = j < lm ? __ldg(row + j) : -1;
            const bool valid = nb >= 0;
            if (!__ballot_sync(FULL, valid)) break;
            // visited test-and-set (scan.rs:366-369): issued now, consumed after the distances -- like the one-shot scan
            // kernel, every listed row is fetched while the (DRAM-resident) bitmap word makes its round trip
            const uint32_t bit = 1u << (nb & 31);
            uint32_t old = 0;
            if (valid) old = atomicOr(bitmap + (nb >> 5), bit);
            const int cnt = __popc(__ballot_sync(FULL, valid));          // valid lanes are exactly 0 .. cnt-1
            if (valid) c.nbr[lane] = nb;
            __syncwarp();
            float dd = 0.0f;
            for (int s = 0; s < cnt; s += c.R) {
                const int nrow = min(c.R, cnt - s);
                gather_rows(c, g, s, nrow);
                if (lane >= s && lane < s + nrow)
                    dd = row_distance<ELEM, METRIC>(c.qsm, c.stage + (size_t)(lane - s) * c.stage_stride, c.nvec);
                __syncwarp();
            }
            const bool fresh = valid && (old & bit) == 0;                // level >= 0 always holds at layer 0
            // a deleted element fails load_element with and without the distance filter (scan.rs:179,376-405):
            // visited, never evaluated, never remembered in `discarded`
            const bool pass = fresh && !(g.flags != nullptr && (__ldg(g.flags + nb) & ELEM_FLAG_DELETED));
            c.n_vis += __popc(__ballot_sync(FULL, fresh));
            c.n_dist += __popc(__ballot_sync(FULL, pass));
            const int32_t my_id = nb;
            const uint32_t dord = f32_ord(dd);
            bool inserted = false;
            uint32_t am = __ballot_sync(FULL, pass && (c.len < c.ef || dord < c.f_ord));
            while (am) {
                const int r = __ffs(am) - 1;
                am &= am - 1;
                const float dr = __shfl_sync(FULL, dd, r);
                const int32_t idr = __shfl_sync(FULL, my_id, r);
                if (c.len < c.ef || f32_ord(dr) < c.f_ord) {                 // scan.rs:371,195-199
                    S.insert(c, make_key(dr, (uint32_t)idr), d);
                    if (lane == r) inserted = true;
                }
            }
            // rejected by distance: remembered for a later resume (scan.rs:385-405; level >= 0 always holds)
            disc_append(d, pass && !inserted, make_key(dd, (uint32_t)my_id), lane);
            __syncwarp();
        }
    }
}

template <int ELEM, int METRIC>
__global__ void __launch_bounds__(SEARCH_THREADS) hnsw_iter_scan_kernel(GraphView g, IterParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cap = p.ef + p.tail_cap;
    const size_t s_bytes = (((size_t)(cap + 1) * 8) + 15) & ~(size_t)15;
    const size_t per_warp = (size_t)p.R * p.stage_stride + g.row_stride + s_bytes + 128 + 16;
    uint8_t *base = smem + (size_t)warp * per_warp;

    WarpCtx c;
    c.stage = base;
    c.qsm = c.stage + (size_t)p.R * p.stage_stride;
    uint8_t *s_mem = c.qsm + g.row_stride;
    c.nbr = reinterpret_cast<int32_t *>(s_mem + s_bytes);
    c.mbar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(c.nbr) + 128);
    IterList S;
    S.S = reinterpret_cast<uint64_t *>(s_mem);
    S.hint = 0;
    SmemList up;                                  // the ef = 1 descent reuses the same list memory
    up.S = S.S;
    up.hint = 0;
    c.stage_stride = p.stage_stride;
    c.R = p.R;
    c.chunk_bytes = c.chunk_stride = 0;
    c.nvec = (int)(g.row_stride / 16);
    c.cap = cap;
    c.lane = lane;
    c.phase = 0;
    c.phase2 = 0;
    c.stage2 = nullptr;
    c.policy = l2_policy_evict_first();
    c.rs_log = nullptr;
    c.rs_cap = c.rs_n = 0;
    c.rs_ent = 3;
    c.spec = true;
    c.hybrid = false;
    const int slot = blockIdx.x * (blockDim.x / 32) + warp;
    c.vis = p.vis_up + (size_t)slot * ITER_UP_ENTRIES;
    if (lane == 0) { mbar_init(c.mbar, 1); fence_mbar_init(); }
    __syncwarp();

    while (true) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.work_counter, 1);
        w = __shfl_sync(FULL, w, 0);
        if (w >= p.nq) break;
        const int qi = w;
        IterQueryState st = p.state[qi];
        c.n_dist = c.n_exp0 = c.n_expu = c.n_vis = 0;
        c.status = 0;
        const bool wanted = p.want_more == nullptr || p.want_more[qi] != 0;
        bool q_loaded = false;
        // p.rounds refills of so.results (scan.rs:817-856) in one launch: amgettuple asks for the next batch only when
        // the previous one is used up, so handing out several batches at once gives the same stream -- a consumer that
        // stops early has merely had the later ones computed for nothing
        for (int rr = 0; rr < p.rounds; rr++) {
        int emitted = 0;
        int32_t *o_ids = p.out_ids + ((size_t)qi * p.rounds + rr) * p.ef;
        float *o_dists = p.out_dists + ((size_t)qi * p.rounds + rr) * p.ef;
        if (wanted && !st.done && c.status == 0) {
            if (!q_loaded) {
                const uint8_t *q = p.queries + (size_t)qi * p.q_stride;
                for (uint32_t i = lane; i < g.row_stride; i += 32) c.qsm[i] = i < p.q_bytes ? __ldg(q + i) : (uint8_t)0;
                q_loaded = true;
            }
            __syncwarp();
            DiscWriter d;
            d.pool = p.pool; d.cmin = p.cmin; d.cfill = p.cfill; d.next = p.next; d.pool_next = p.pool_next; d.pool_cap = p.pool_cap;
            d.head = st.head; d.tail = st.tail; d.fill = st.fill; d.live = st.live; d.status = 0;
            d.tail_min = st.tail >= 0 ? p.cmin[st.tail] : KEY_NONE;
            uint32_t *bitmap = p.bitmap + (size_t)qi * p.n_words;
            int n_res = 0;
            if (st.round == 0) {
                // scan.rs:470-472; a deleted entry point ends the scan the same way (scan.rs:475-486 -> :179)
                if (g.entry >= 0 && !(g.flags != nullptr && (__ldg(g.flags + g.entry) & ELEM_FLAG_DELETED))) {
                    if (lane == 0) c.nbr[0] = g.entry;
                    __syncwarp();
                    gather_rows(c, g, 0, 1);
                    float d0 = 0.0f;
                    if (lane == 0) d0 = row_distance<ELEM, METRIC>(c.qsm, c.stage, c.nvec);
                    d0 = __shfl_sync(FULL, d0, 0);
                    c.n_dist = 1;
                    __syncwarp();
                    uint64_t ep_key = make_key(d0, (uint32_t)g.entry);
                    c.ef = 1;
                    for (int lc = g.entry_level; lc >= 1 && c.status == 0; lc--) {   // scan.rs:492-512
                        vis_clear(c.vis, ITER_UP_ENTRIES, lane);
                        if (lane == 0) vis_insert(c.vis, ITER_UP_ENTRIES - 1, key_id(ep_key));
                        __syncwarp();
                        c.n_vis++;
                        up.reset1(c, ep_key);
                        search_layer<ELEM, METRIC, true>(c, up, g, lc, ITER_UP_ENTRIES - 1, ITER_UP_ENTRIES / 4 * 3 - 32);
                        ep_key = up.get(c, 0) & ~1ull;
                    }
                    if (c.status == 0) {                                         // scan.rs:515-528
                        const uint32_t eid = key_id(ep_key);
                        if (lane == 0) atomicOr(bitmap + (eid >> 5), 1u << (eid & 31));
                        __syncwarp();
                        c.n_vis++;
                        c.ef = p.ef;
                        if (lane == 0) S.S[0] = ep_key;
                        __syncwarp();
                        c.len = 1;
                        S.hint = 0;
                        if (c.ef == 1) c.f_ord = key_ord(ep_key);
                        search_layer_iter<ELEM, METRIC>(c, S, g, bitmap, d);
                        n_res = min(c.len, c.ef);
                    }
                }
            } else if (p.mode != 0 && d.live > 0) {
                c.ef = p.ef;
                const int n_sel = disc_pop_nearest(d, S.S, p.ef, lane);          // scan.rs:552-561 / :831-841
                c.len = n_sel;
                S.hint = 0;
                if (st.tuples < p.max_scan_tuples) {
                    if (c.len >= c.ef) c.f_ord = key_ord(S.S[c.ef - 1]);
                    search_layer_iter<ELEM, METRIC>(c, S, g, bitmap, d);         // scan.rs:563-576
                }
                n_res = min(c.len, c.ef);
            }
            c.status |= d.status;
            // ---- emission (scan.rs:794-875): nearest first; strict_order drops what would go backwards ----
            int n_counted = n_res;
            if (g.flags == nullptr) {
                int skip = 0;
                if (p.mode == 2) {
                    for (int b0 = 0; b0 < n_res; b0 += 32) {
                        const int i = b0 + lane;
                        skip += __popc(__ballot_sync(FULL, i < n_res && key_ord(S.S[i]) < st.prev_ord));
                    }
                }
                emitted = n_res - skip;
                for (int i = lane; i < emitted; i += 32) {
                    const uint64_t key = S.S[skip + i];
                    o_ids[i] = (int32_t)key_id(key);
                    o_dists[i] = ord_f32(key_ord(key));
                }
                if (emitted > 0) st.prev_ord = key_ord(S.S[n_res - 1]);
            } else {
                // vacuumed index: an element without heap TIDs is skipped before it counts as a tuple and before the
                // strict-order test (scan.rs:866-871 precede :801-806)
                uint32_t last_ord = st.prev_ord;
                for (int b0 = 0; b0 < n_res; b0 += 32) {
                    const int i = b0 + lane;
                    uint64_t key = 0;
                    bool has_tid = false;
                    if (i < n_res) {
                        key = S.S[i];
                        has_tid = !(__ldg(g.flags + key_id(key)) & ELEM_FLAG_NO_HEAPTIDS);
                    }
                    n_counted -= __popc(__ballot_sync(FULL, i < n_res && !has_tid));
                    const bool out = has_tid && !(p.mode == 2 && key_ord(key) < st.prev_ord);
                    const uint32_t om = __ballot_sync(FULL, out);
                    if (out) {
                        const int pos = emitted + __popc(om & ((1u << lane) - 1u));
                        o_ids[pos] = (int32_t)key_id(key);
                        o_dists[pos] = ord_f32(key_ord(key));
                    }
                    if (om) last_ord = __shfl_sync(FULL, key_ord(key), 31 - __clz(om));   // ascending: the last one out is the largest
                    emitted += __popc(om);
                }
                st.prev_ord = last_ord;
            }
            st.tuples += n_counted;                                              // scan.rs:871
            st.round += 1;
            disc_flush_tail(d, lane);
            st.head = d.head; st.tail = d.tail; st.fill = d.fill; st.live = d.live;
            if (n_res == 0 || p.mode == 0 || d.live == 0) st.done = 1;            // nothing left to resume from
            __syncwarp();
        }
        for (int i = emitted + lane; i < p.ef; i += 32) {
            o_ids[i] = -1;
            o_dists[i] = __int_as_float(0x7f800000);
        }
        if (lane == 0) p.out_counts[(size_t)qi * p.rounds + rr] = emitted;
        __syncwarp();
        }   // rounds
        if (lane == 0) {
            p.state[qi] = st;
            p.out_done[qi] = (uint8_t)st.done;
            p.out_status[qi] = c.status;
            if (p.out_ctr) reinterpret_cast<uint4 *>(p.out_ctr)[qi] = make_uint4(c.n_dist, c.n_exp0, c.n_expu, c.n_vis);
        }
        __syncwarp();
    }
}

typedef void (*iter_kernel_t)(GraphView, IterParams);

static iter_kernel_t pick_iter(int elem, int metric) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return hnsw_iter_scan_kernel<ELEM_F32, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return hnsw_iter_scan_kernel<ELEM_F32, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return hnsw_iter_scan_kernel<ELEM_F32, METRIC_L1>;
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return hnsw_iter_scan_kernel<ELEM_F16, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return hnsw_iter_scan_kernel<ELEM_F16, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return hnsw_iter_scan_kernel<ELEM_F16, METRIC_L1>;
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return hnsw_iter_scan_kernel<ELEM_BIT, METRIC_HAMMING>;
        if (metric == METRIC_JACCARD) return hnsw_iter_scan_kernel<ELEM_BIT, METRIC_JACCARD>;
    }
    return nullptr;
}

size_t iter_warp_smem(int R, uint32_t stage_stride, uint32_t row_stride, int ef, int tail_cap) {
    return (size_t)R * stage_stride + row_stride + ((((size_t)(ef + tail_cap + 1) * 8) + 15) & ~(size_t)15) + 128 + 16;
}

cudaError_t iter_kernel_occupancy(int elem, int metric, size_t smem_block, int *blocks_per_sm) {
    iter_kernel_t kfn = pick_iter(elem, metric);
    if (!kfn) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_block);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kfn, ITER_WARPS * 32, smem_block);
}

cudaError_t launch_iter_scan(int elem, int metric, const GraphView &g, const IterParams &p, int grid, size_t smem_block,
                             cudaStream_t stream) {
    iter_kernel_t kfn = pick_iter(elem, metric);
    if (!kfn) return cudaErrorInvalidValue;
    kfn<<<grid, ITER_WARPS * 32, smem_block, stream>>>(g, p);
    return cudaGetLastError();
}

}  // namespace pgvrx
