// build_host.cu -- pgvrx_index_build: the reference's ambuild insert loop (src/index/build.rs:400-535)
// with the search_layer calls of find_element_neighbors (src/graph/mod.rs:355-427) batched on the GPU.
//
// Division of labour (north star #2): the GPU scores each pending element's ef_construction
// candidates (K2, hnsw_build.cu); the HOST applies select_neighbors, the duplicate rule,
// update_neighbor_connections and the entry-point rule in the reference's order, so every neighbour
// list comes out identical to a sequential build.
//
// Exactness under batching.  A round searches B pending elements against the committed snapshot and commits them in
// insert order.  The kernel logs every expansion of a search: node, layer, W's furthest distance f_t at that moment
// (or "W not full") and, per entry of the list it read, whether that entry is nearer than f_t.  Element j's
// speculative result is committed only if, for every logged expansion whose list has changed since the snapshot, the
// entries nearer than f_t form the same sequence in the old and in the current list (with W not full: the list is
// unchanged): only those can be accepted then or later, everything else merely toggles the visited flag of an
// element that is rejected wherever it is met, so a sequential run would have replayed the very same search.
// Otherwise -- or when the entry point moved or a duplicate was merged -- the round stops before j and j is searched
// again next round (the head of a round is always valid, so progress is guaranteed).
//
// select_neighbors never needs a distance, only "d(e, r) <= d(e, owner)" for kept r (graph/mod.rs:333): the GPU
// answers that for all pairs of a group at once (pairwise_kernel, 128-bit masks), for the new elements' candidate
// lists and for every list that pending elements link back to.  Host distances remain only for groups that do not
// fit on a CTA (ef_construction > 128 or very wide rows); the pending-vs-pending distances the validation needs for
// newcomers of a changed list come from the GPU as well (pending_pairs_kernel).  For the bench configuration no
// distance is evaluated on the host (pgvrx_build_stats.n_host_dist == 0).
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_fp16.h>

#include "index_internal.h"

using namespace pgvrx;

namespace pgvrx_host {

// ------------------------------------------------------------------------------------------------
// strict-order host distances (a13-a15), returned as the build path stores them: f64 -> f32
// (src/index/build.rs:346-369).  Sequential accumulation, no contraction (nvcc host flags have no
// -ffast-math and this file is built with -ffp-contract=off via -Xcompiler).
// ------------------------------------------------------------------------------------------------
static inline float h2f(uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); }

static float host_distance(int elem, int metric, int dim, const uint8_t *pa, const uint8_t *pb) {
    if (elem == PGVRX_ELEM_F32) {
        const float *a = (const float *)pa, *b = (const float *)pb;
        float acc = 0.0f;
        if (metric == PGVRX_METRIC_L2SQ) { for (int i = 0; i < dim; i++) { float d = a[i] - b[i]; acc += d * d; } return acc; }
        if (metric == PGVRX_METRIC_NEG_IP) { for (int i = 0; i < dim; i++) acc += a[i] * b[i]; return (float)(-(double)acc); }
        for (int i = 0; i < dim; i++) acc += std::fabs(a[i] - b[i]);
        return acc;
    }
    if (elem == PGVRX_ELEM_F16) {
        const uint16_t *a = (const uint16_t *)pa, *b = (const uint16_t *)pb;
        float acc = 0.0f;
        if (metric == PGVRX_METRIC_L2SQ) { for (int i = 0; i < dim; i++) { float d = h2f(a[i]) - h2f(b[i]); acc += d * d; } return acc; }
        if (metric == PGVRX_METRIC_NEG_IP) { for (int i = 0; i < dim; i++) acc += h2f(a[i]) * h2f(b[i]); return (float)(-(double)acc); }
        for (int i = 0; i < dim; i++) acc += std::fabs(h2f(a[i]) - h2f(b[i]));
        return acc;
    }
    const size_t nb = ((size_t)dim + 7) / 8;
    if (metric == PGVRX_METRIC_HAMMING) {
        uint64_t d = 0;
        for (size_t i = 0; i < nb; i++) d += (uint64_t)__builtin_popcount((unsigned)(pa[i] ^ pb[i]));
        return (float)(double)d;
    }
    uint64_t ab = 0, aa = 0, bb = 0;
    for (size_t i = 0; i < nb; i++) {
        ab += (uint64_t)__builtin_popcount((unsigned)(pa[i] & pb[i]));
        aa += (uint64_t)__builtin_popcount((unsigned)pa[i]);
        bb += (uint64_t)__builtin_popcount((unsigned)pb[i]);
    }
    if (ab == 0) return 1.0f;
    return (float)(1.0 - ((double)ab / (double)(aa + bb - ab)));
}

// ------------------------------------------------------------------------------------------------
// tiny fork-join pool
// ------------------------------------------------------------------------------------------------
// Workers spin on a generation counter while a build is running: a round issues several short parallel sections
// (tens of microseconds each), and waking sleeping threads through a condition variable would cost more than the work.
class Pool {
  public:
    explicit Pool(int n) : n_(std::max(1, n)) {
        for (int t = 1; t < n_; t++) th_.emplace_back([this] { worker(); });
    }
    ~Pool() {
        stop_.store(true, std::memory_order_release);
        gen_.fetch_add(1, std::memory_order_release);
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    // run fn(i) for i in [0, count) on all threads (dynamic distribution)
    void parallel_for(int64_t count, const std::function<void(int64_t)> &fn) {
        if (count <= 0) return;
        if (n_ == 1 || count == 1) { for (int64_t i = 0; i < count; i++) fn(i); return; }
        fn_ = &fn;
        count_ = count;
        next_.store(0, std::memory_order_relaxed);
        pending_.store(n_ - 1, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        run();
        for (int spins = 0; pending_.load(std::memory_order_acquire) != 0; spins++) {
            cpu_relax();
            if (spins > 4000) std::this_thread::yield();
        }
    }

  private:
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void run() {
        // small items (thousands of groups) are handed out in chunks so that the shared counter is not the bottleneck
        const int64_t chunk = std::max<int64_t>(1, count_ / ((int64_t)n_ * 8));
        for (;;) {
            const int64_t i0 = next_.fetch_add(chunk, std::memory_order_relaxed);
            if (i0 >= count_) break;
            const int64_t i1 = std::min(count_, i0 + chunk);
            for (int64_t i = i0; i < i1; i++) (*fn_)(i);
        }
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            uint64_t gnow;
            int spins = 0;
            while ((gnow = gen_.load(std::memory_order_acquire)) == seen) {
                cpu_relax();
                ++spins;
                if (spins > 200000) { std::this_thread::sleep_for(std::chrono::microseconds(50)); spins = 100000; }   // long GPU phase: back off
                else if (spins > 4000) std::this_thread::yield();   // never starve the main thread when cores are oversubscribed
            }
            seen = gnow;
            if (stop_.load(std::memory_order_acquire)) return;
            run();
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    int n_;
    std::vector<std::thread> th_;
    const std::function<void(int64_t)> *fn_ = nullptr;
    int64_t count_ = 0;
    std::atomic<int64_t> next_{0};
    std::atomic<int> pending_{0};
    std::atomic<uint64_t> gen_{0};
    std::atomic<bool> stop_{false};
};

// ------------------------------------------------------------------------------------------------
// host mirror of the graph (graph/mod.rs:26-84: neighbour lists keep the distance to their owner)
// ------------------------------------------------------------------------------------------------
struct Cand { float distance; int32_t idx; };

struct BuildState {
    int elem = 0, metric = 0, dim = 0, m = 0, efc = 0;
    size_t row_bytes = 0;
    int64_t n = 0, n_upper = 0, entry = -1;
    int entry_level = -1;
    std::vector<uint8_t> values;        // rows of kept elements, element id * row_bytes (build.rs:242)
    std::vector<uint8_t> level;
    std::vector<int32_t> upper_base;
    std::vector<int32_t> adj0;          // n x 2m
    std::vector<float> dist0;
    std::vector<uint8_t> cnt0;
    std::vector<int32_t> adj_up;        // n_upper x m
    std::vector<float> dist_up;
    std::vector<uint8_t> cnt_up;
    std::vector<uint64_t> tids;         // n x 10 (build.rs:245)
    std::vector<uint8_t> tid_cnt;
    std::vector<uint8_t> del;           // HnswElementTupleData.deleted (set by pgvrx_index_vacuum's last pass)
    std::vector<uint32_t> stamp;        // per node: round in which one of its lists was last modified
    uint32_t round = 0;
    std::atomic<uint64_t> n_pair{0};
    struct RoundScratch *scratch = nullptr;   // device / pinned round buffers, kept across appending calls

    const uint8_t *row(int64_t e) const { return values.data() + (size_t)e * row_bytes; }
    std::atomic<uint64_t> n_host_dist{0};   // distances evaluated on the host (groups that do not fit on a CTA)
    float dist(int64_t a, int64_t b) { n_host_dist.fetch_add(1, std::memory_order_relaxed); return host_distance(elem, metric, dim, row(a), row(b)); }
    int lm(int lc) const { return lc == 0 ? 2 * m : m; }
    int32_t *ids(int64_t v, int lc) { return lc == 0 ? &adj0[(size_t)v * 2 * m] : &adj_up[(size_t)(upper_base[(size_t)v] + lc - 1) * m]; }
    float *dists(int64_t v, int lc) { return lc == 0 ? &dist0[(size_t)v * 2 * m] : &dist_up[(size_t)(upper_base[(size_t)v] + lc - 1) * m]; }
    uint8_t &cnt(int64_t v, int lc) { return lc == 0 ? cnt0[(size_t)v] : cnt_up[(size_t)(upper_base[(size_t)v] + lc - 1)]; }
};

void destroy_build_state(BuildState *b);   // defined below, after RoundScratch

// a3: select_neighbors + check_element_closer (graph/mod.rs:269-339)
static void select_neighbors(BuildState &g, const Cand *cand, int n_cand, int max_neighbors, std::vector<Cand> &out) {
    out.clear();
    if (n_cand <= max_neighbors) { out.assign(cand, cand + n_cand); return; }     // :276
    std::vector<Cand> discarded;
    uint64_t pairs = 0;
    for (int i = 0; i < n_cand; i++) {
        if ((int)out.size() >= max_neighbors) break;                               // :285
        const Cand &e = cand[i];
        bool closer = true;
        for (const Cand &r : out) {                                                // :324-336
            float d = g.dist(e.idx, r.idx);
            pairs++;
            if (d <= e.distance) { closer = false; break; }
        }
        if (closer) out.push_back(e); else discarded.push_back(e);
    }
    for (const Cand &d : discarded) {                                              // :300-305
        if ((int)out.size() >= max_neighbors) break;
        out.push_back(d);
    }
    g.n_pair.fetch_add(pairs, std::memory_order_relaxed);
}

struct Pending {                 // one element of the current round, after phase A
    int64_t input_row;
    int level;
    std::vector<std::vector<Cand>> sel;   // [lc] selected neighbours (graph/mod.rs:419-422)
    int dup_of = -1;                      // element id this row merges into (build.rs:482-503)
};

// one (node, layer) list that pending elements of the current round link back to
struct Group {
    int32_t node, lc;
    int n_list;                       // list length at the start of the round
    int row_off;                      // first row of the group in the GPU mask arrays, -1: distances on the host
    std::vector<int32_t> ids;         // the list as the round's searches saw it, then the pending elements linking to it
    std::vector<float> own;           // distance of every entry of `ids` to the node
    // the list after each of the pending elements has been linked in, in order (graph/mod.rs:458-487), computed for
    // the whole round up front -- a group's evolution depends on nothing but the group -- so that the sequential commit
    // only advances `step`: state k (0-based) = list after members 0..k, lm entries each
    std::vector<int32_t> st_ids;
    std::vector<float> st_ds;
    std::vector<int16_t> st_pos;      // position in `ids` of every entry of that list
    std::vector<uint8_t> st_cnt;
    int lm;
    int step;                         // members committed so far (0: the list is still the one the searches saw)
    bool modified;
};

static int16_t g_identity_pos[256];

// the list of a group as it stands after `step` committed members
static inline void group_list(const Group &G, const int32_t *&ids, const float *&ds, const int16_t *&pos, int &cnt) {
    if (G.step == 0) { ids = G.ids.data(); ds = G.own.data(); pos = g_identity_pos; cnt = G.n_list; return; }
    const size_t o = (size_t)(G.step - 1) * G.lm;
    ids = &G.st_ids[o]; ds = &G.st_ds[o]; pos = &G.st_pos[o]; cnt = G.st_cnt[(size_t)G.step - 1];
}

// order-preserving bits of a distance, as the kernels compare them (hnsw_device.cuh f32_ord)
static inline uint32_t ord_of(float d) {
    d = d + 0.0f;
    uint32_t b;
    memcpy(&b, &d, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// select_neighbors (graph/mod.rs:269-308) with check_element_closer answered by the GPU's masks:
// masks[pos * W .. +W) has bit r set iff d(member pos, member r) <= own[pos]  (W = PAIR_MAX_ROWS / 32 words)
static void select_neighbors_mask(BuildState &g, const Cand *cand, const int16_t *pos, int n_cand, int max_neighbors,
                                  const uint32_t *masks, std::vector<Cand> &out, std::vector<int16_t> &out_pos) {
    constexpr int W = PAIR_MAX_ROWS / 32;
    out.clear();
    out_pos.clear();
    if (n_cand <= max_neighbors) { out.assign(cand, cand + n_cand); out_pos.assign(pos, pos + n_cand); return; }
    uint32_t kept[W] = {0};
    int discarded[PAIR_MAX_ROWS + 1];
    int n_disc = 0;
    uint64_t pairs = 0;
    for (int i = 0; i < n_cand; i++) {
        if ((int)out.size() >= max_neighbors) break;
        const uint32_t *c = masks + (size_t)pos[i] * W;
        uint32_t hit = 0;
        for (int w = 0; w < W; w++) hit |= c[w] & kept[w];
        pairs += out.size();
        if (!hit) {
            out.push_back(cand[i]);
            out_pos.push_back(pos[i]);
            kept[pos[i] >> 5] |= 1u << (pos[i] & 31);
        } else {
            discarded[n_disc++] = i;
        }
    }
    for (int t = 0; t < n_disc; t++) {
        if ((int)out.size() >= max_neighbors) break;
        out.push_back(cand[discarded[t]]);
        out_pos.push_back(pos[discarded[t]]);
    }
    g.n_pair.fetch_add(pairs, std::memory_order_relaxed);
}

// update_neighbor_connections for one (node, layer) and all pending elements that link to it, in insert order
// (graph/mod.rs:458-487): append while there is room, otherwise stable-sort list + newcomer by distance and re-select.
static void simulate_group(BuildState &g, Group &G, const uint32_t *masks /* rows of this group or NULL */) {
    const int lm = G.lm, n_members = (int)G.ids.size() - G.n_list;
    G.st_ids.assign((size_t)n_members * lm, -1);
    G.st_ds.assign((size_t)n_members * lm, 0.0f);
    G.st_pos.assign((size_t)n_members * lm, -1);
    G.st_cnt.assign((size_t)n_members, 0);
    Cand cur[256], all[257];
    int16_t cur_pos[256], all_pos[257];
    int order[257];
    int cnt = G.n_list;
    for (int t = 0; t < cnt; t++) { cur[t] = Cand{G.own[(size_t)t], G.ids[(size_t)t]}; cur_pos[t] = (int16_t)t; }
    std::vector<Cand> sel_out;
    std::vector<int16_t> sel_pos;
    for (int k = 0; k < n_members; k++) {
        const int16_t epos = (int16_t)(G.n_list + k);
        const Cand nw{G.own[(size_t)epos], G.ids[(size_t)epos]};
        if (cnt < lm) {                                          // :469-471 room available
            cur[cnt] = nw;
            cur_pos[cnt] = epos;
            cnt++;
        } else {                                                 // :473-485
            const int n_all = cnt + 1;
            for (int t = 0; t < n_all; t++) order[t] = t;
            auto dist_of = [&](int t) { return t < cnt ? cur[t].distance : nw.distance; };
            std::stable_sort(order, order + n_all, [&](int x, int y) { return dist_of(x) < dist_of(y); });
            for (int t = 0; t < n_all; t++) {
                const int src = order[t];
                all[t] = src < cnt ? cur[src] : nw;
                all_pos[t] = src < cnt ? cur_pos[src] : epos;
            }
            if (masks) select_neighbors_mask(g, all, all_pos, n_all, lm, masks, sel_out, sel_pos);
            else { select_neighbors(g, all, n_all, lm, sel_out); sel_pos.resize(sel_out.size()); }
            cnt = (int)sel_out.size();
            for (int t = 0; t < cnt; t++) { cur[t] = sel_out[(size_t)t]; cur_pos[t] = masks ? sel_pos[(size_t)t] : (int16_t)-1; }
            if (!masks) {
                // positions are still needed by the validation: recover them by id (small groups only take this path
                // when they do not fit on a CTA, i.e. rarely)
                for (int t = 0; t < cnt; t++)
                    for (size_t u = 0; u < G.ids.size(); u++) if (G.ids[u] == cur[t].idx) { cur_pos[t] = (int16_t)u; break; }
            }
        }
        const size_t o = (size_t)k * lm;
        for (int t = 0; t < cnt; t++) { G.st_ids[o + t] = cur[t].idx; G.st_ds[o + t] = cur[t].distance; G.st_pos[o + t] = cur_pos[t]; }
        G.st_cnt[(size_t)k] = (uint8_t)cnt;
    }
}

// aminsert's back-link rule for one (node, layer) and all pending elements that link to it, in insert order
// (src/index/insert.rs:500-738 get_update_index + :793-870 write_neighbor_update): append while a slot is free;
// otherwise sort the list by distance to the node (stable: ties keep slot order, the new element last among equals),
// run the closer-test heuristic in which the new element is never pruned and never prunes, and if the new element is
// among the lm kept, overwrite the first slot (in stored order) whose element was not kept.  Slots keep their places.
// On an index a vacuum has touched (`flagged`): a full list that still holds an element flagged deleted, or one left
// without heap TIDs, gives that slot -- the first such in stored order -- to the new element at once (:569-623).
// Elements at or above `e_base` are the pending ones of this round (their heap TID is assigned when they commit).
static void simulate_group_disk(BuildState &g, Group &G, const uint32_t *masks /* rows of this group or NULL */,
                                const bool flagged = false, const int64_t e_base = 0) {
    constexpr int W = PAIR_MAX_ROWS / 32;
    const int lm = G.lm, n_members = (int)G.ids.size() - G.n_list;
    G.st_ids.assign((size_t)n_members * lm, -1);
    G.st_ds.assign((size_t)n_members * lm, 0.0f);
    G.st_pos.assign((size_t)n_members * lm, -1);
    G.st_cnt.assign((size_t)n_members, 0);
    Cand cur[256];
    int16_t cur_pos[256];
    int order[256], pruned[256];
    bool kept_slot[256];
    int cnt = G.n_list;
    for (int t = 0; t < cnt; t++) { cur[t] = Cand{G.own[(size_t)t], G.ids[(size_t)t]}; cur_pos[t] = (int16_t)t; }
    uint64_t pairs = 0;
    for (int k = 0; k < n_members; k++) {
        const int16_t epos = (int16_t)(G.n_list + k);
        const Cand nw{G.own[(size_t)epos], G.ids[(size_t)epos]};
        if (cnt < lm) {                                          // insert.rs:559-562 -> :843-853 first free slot
            cur[cnt] = nw;
            cur_pos[cnt] = epos;
            cnt++;
        } else if (flagged && [&]() {
                       for (int t = 0; t < cnt; t++) {
                           const int64_t x = cur[t].idx;
                           if (x < e_base && (g.del[(size_t)x] || g.tid_cnt[(size_t)x] == 0)) { cur[t] = nw; cur_pos[t] = epos; return true; }
                       }
                       return false;
                   }()) {
            // (slot of a deleted / emptied element reused, insert.rs:620-623)
        } else {
            for (int t = 0; t < cnt; t++) { order[t] = t; kept_slot[t] = false; }
            std::stable_sort(order, order + cnt, [&](int x, int y) { return cur[x].distance < cur[y].distance; });   // :628-632
            // the new element sorts after every existing entry that is not farther (:652-663)
            int n_before = 0;
            while (n_before < cnt && !(nw.distance < cur[order[n_before]].distance)) n_before++;
            int n_sel = 0, n_pruned = 0;
            bool new_selected = false;
            uint32_t kept[W] = {0};
            int kept_list[256], n_kept = 0;
            for (int t = 0; t <= cnt && n_sel < lm; t++) {       // :671-705
                if (t == n_before) { new_selected = true; n_sel++; if (n_sel >= lm) break; }
                if (t == cnt) break;
                const int slot = order[t];
                bool closer = true;
                if (masks) {
                    const uint32_t *c = masks + (size_t)cur_pos[slot] * W;
                    uint32_t hit = 0;
                    for (int w = 0; w < W; w++) hit |= c[w] & kept[w];
                    pairs += (uint64_t)n_kept;
                    closer = hit == 0;
                } else {
                    for (int u = 0; u < n_kept; u++) {
                        pairs++;
                        if (g.dist(cur[slot].idx, cur[kept_list[u]].idx) <= cur[slot].distance) { closer = false; break; }
                    }
                }
                if (closer) {
                    kept_slot[slot] = true;
                    kept[cur_pos[slot] >> 5] |= 1u << (cur_pos[slot] & 31);
                    kept_list[n_kept++] = slot;
                    n_sel++;
                } else {
                    pruned[n_pruned++] = slot;
                }
            }
            for (int t = 0; t < n_pruned && n_sel < lm; t++) { kept_slot[pruned[t]] = true; n_sel++; }   // :708-713
            if (new_selected) {                                  // :716-738
                for (int t = 0; t < cnt; t++)
                    if (!kept_slot[t]) { cur[t] = nw; cur_pos[t] = epos; break; }
            }
        }
        const size_t o = (size_t)k * lm;
        for (int t = 0; t < cnt; t++) { G.st_ids[o + t] = cur[t].idx; G.st_ds[o + t] = cur[t].distance; G.st_pos[o + t] = cur_pos[t]; }
        G.st_cnt[(size_t)k] = (uint8_t)cnt;
    }
    g.n_pair.fetch_add(pairs, std::memory_order_relaxed);
}

}  // namespace pgvrx_host

using namespace pgvrx_host;

namespace {

struct RoundBuffers {
    DevBuf elem_ids, out_pack, vis, work, sc_off, sc_rows, g_ids, g_off, g_own, g_mat;
    void release() {
        DevBuf *b[] = {&elem_ids, &out_pack, &vis, &work, &sc_off, &sc_rows, &g_ids, &g_off, &g_own, &g_mat};
        for (DevBuf *x : b) x->release();
    }
};

// pinned host buffer (device-to-host copies of per-round results run at full PCIe rate and truly asynchronously)
struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 2 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

namespace pgvrx_host {
// cudaMalloc / cudaFree / cudaMallocHost cost tens to hundreds of milliseconds per insert call when every call grows
// its buffers from nothing (measured: 50-250 ms per 6.7k-row call), so they live with the index
struct RoundScratch {
    RoundBuffers rb;
    PinnedBuf h_pack, h_mat;
    // Host-to-device traffic of a round (the pending elements' rows, committed neighbour rows, group descriptions) leaves
    // from page-locked staging: a cudaMemcpyAsync from pageable memory is staged by the driver synchronously (tens of
    // microseconds each, a dozen per round).  Two arenas alternate by round: what round r queued behind its last kernel
    // has certainly been read when round r + 2 starts writing (round r + 1 waited for its own search results in between).
    PinnedBuf h_up[2];
    size_t up_off = 0;
    int up_sel = 0;
    ~RoundScratch() { rb.release(); }
};
void destroy_build_state(BuildState *b) {
    if (b) delete b->scratch;
    delete b;
}
}  // namespace pgvrx_host

extern "C" {

static void reserve_mirror(BuildState &g, int64_t cap) {
    const size_t c = (size_t)std::max<int64_t>(cap, g.n), l0 = (size_t)2 * g.m;
    if (g.values.capacity() >= c * g.row_bytes && g.adj0.capacity() >= c * l0) return;
    try {
        g.values.reserve(c * g.row_bytes);
        g.level.reserve(c);
        g.upper_base.reserve(c);
        g.adj0.reserve(c * l0);
        g.dist0.reserve(c * l0);
        g.cnt0.reserve(c);
        g.tids.reserve(c * PGVRX_HEAPTIDS);
        g.tid_cnt.reserve(c);
        g.del.reserve(c);
        g.stamp.reserve(c);
    } catch (const std::bad_alloc &) {
        // a capacity far beyond what is ever inserted: fall back to the vectors' own geometric growth
    }
}

// Geometry of the CTA-per-search kernel (hnsw_team.cu) for a round of B searches: as many helper warps as still let every
// search of the round be resident at once (a round ends with its slowest search), shared-memory visited table of four
// times the expected visited count (clamped to what fits).  Returns false when nothing fits (the one-warp kernel runs).
static bool team_config(pgvrx_index *ix, const ScanConfig &cfg, int efc, int B, TeamGeom *out, size_t *smem_out, int *grid_out) {
    const int m = ix->desc.m;
    TeamGeom t;
    t.lmp = (2 * m + 31) / 32 * 32;
    t.stage_bytes = cfg.L.stage_bytes;
    t.row_stride = ix->row_stride;
    t.list_j = cfg.L.list_j;
    t.ef = efc;
    t.tail_cap = cfg.L.tail_cap;
    const uint32_t vis_want = std::min<uint32_t>(16384u, std::max<uint32_t>(2048u, pgvrx_pow2_at_least((uint64_t)efc * 2 * m * 4)));
    int forced = 0;
    if (const char *e = getenv("PGVRX_TEAM_WARPS")) forced = atoi(e);
    static const int options[3] = {8, 4, 2};
    for (int oi = 0; oi < 3; oi++) {
        t.warps = forced >= 2 && forced <= TEAM_MAX_WARPS ? forced : options[oi];
        t.vis_entries = vis_want;
        t.n_slots = std::min(TEAM_MAX_SLOTS, std::max(4, 4 * t.warps));
        size_t smem = team_smem_bytes(t);
        while (smem > 220 * 1024 && t.n_slots > 8) { t.n_slots -= 4; smem = team_smem_bytes(t); }
        while (smem > 220 * 1024 && t.vis_entries > 4096) { t.vis_entries /= 2; smem = team_smem_bytes(t); }
        if (smem > 220 * 1024) { if (forced) return false; continue; }
        int occ = 0;
        if (team_kernel_occupancy(ix->desc.elem_type, ix->desc.metric, t.list_j, smem, t.warps * 32, &occ) != cudaSuccess || occ < 1) {
            cudaGetLastError();
            if (forced) return false;
            continue;
        }
        if ((int64_t)ix->sm_count * occ >= B || oi == 2 || forced) {
            *out = t;
            *smem_out = smem;
            *grid_out = (int)std::min<int64_t>(B, (int64_t)ix->sm_count * occ);
            return true;
        }
    }
    return false;
}

// Host mirror of a graph that came in through pgvrx_index_load_graph: arrays back from the device, edge distances
// recomputed (the on-disk format stores none; get_update_index recomputes them per insert, insert.rs:569-617).
static int mirror_from_device(pgvrx_index *ix, int nthreads) {
    const int m = ix->desc.m, l0 = 2 * m;
    BuildState *b = new BuildState();
    b->elem = ix->desc.elem_type; b->metric = ix->desc.metric; b->dim = ix->desc.dim; b->m = m; b->efc = ix->desc.ef_construction;
    b->row_bytes = ix->row_bytes;
    struct Guard { BuildState *b; bool ok; ~Guard() { if (!ok) destroy_build_state(b); } } guard{b, false};   // a failed copy leaves no half-built mirror
    BuildState &g = *b;
    const int64_t n = ix->n, nu = ix->n_upper;
    g.n = n; g.n_upper = nu; g.entry = ix->entry; g.entry_level = ix->entry_level;
    reserve_mirror(g, ix->desc.capacity);
    g.values.resize((size_t)n * g.row_bytes);
    g.level.resize((size_t)n);
    g.upper_base.resize((size_t)n);
    g.adj0.resize((size_t)n * l0, -1);
    g.dist0.resize((size_t)n * l0, 0.0f);
    g.cnt0.resize((size_t)n, 0);
    g.adj_up.resize((size_t)nu * m, -1);
    g.dist_up.resize((size_t)nu * m, 0.0f);
    g.cnt_up.resize((size_t)nu, 0);
    g.tids.assign((size_t)n * PGVRX_HEAPTIDS, ~0ull);
    g.tid_cnt.assign((size_t)n, 1);
    g.del.assign((size_t)n, 0);
    g.stamp.assign((size_t)n, 0);
    if (n == 0) { guard.ok = true; ix->build = b; return PGVRX_OK; }
    CU_TRY(ix, cudaMemcpy2D(g.values.data(), g.row_bytes, ix->vectors.p, ix->row_stride, g.row_bytes, (size_t)n, cudaMemcpyDeviceToHost));
    CU_TRY(ix, cudaMemcpy(g.level.data(), ix->level.p, (size_t)n, cudaMemcpyDeviceToHost));
    CU_TRY(ix, cudaMemcpy(g.upper_base.data(), ix->upper_base.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CU_TRY(ix, cudaMemcpy(g.adj0.data(), ix->adj0.p, (size_t)n * l0 * 4, cudaMemcpyDeviceToHost));
    if (nu > 0) CU_TRY(ix, cudaMemcpy(g.adj_up.data(), ix->adj_up.p, (size_t)nu * m * 4, cudaMemcpyDeviceToHost));
    for (int64_t e = 0; e < n; e++) g.tids[(size_t)e * PGVRX_HEAPTIDS] = (uint64_t)e;
    if (ix->has_flags) {                 // an index a vacuum has touched before: deleted / emptied elements
        std::vector<uint8_t> f((size_t)n);
        CU_TRY(ix, cudaMemcpy(f.data(), ix->flags.p, (size_t)n, cudaMemcpyDeviceToHost));
        for (int64_t e = 0; e < n; e++) {
            g.del[(size_t)e] = (f[(size_t)e] & PGVRX_ELEM_DELETED) ? 1 : 0;
            if (f[(size_t)e] & PGVRX_ELEM_NO_HEAPTIDS) { g.tid_cnt[(size_t)e] = 0; g.tids[(size_t)e * PGVRX_HEAPTIDS] = ~0ull; }
        }
    }
    Pool pool(nthreads);
    pool.parallel_for(n, [&](int64_t v) {
        for (int lc = 0; lc <= g.level[(size_t)v]; lc++) {
            const int lm = g.lm(lc);
            int32_t *ids = g.ids(v, lc);
            float *ds = g.dists(v, lc);
            int c = 0;
            while (c < lm && ids[c] >= 0) { ds[c] = host_distance(g.elem, g.metric, g.dim, g.row(v), g.row(ids[c])); c++; }
            g.cnt(v, lc) = (uint8_t)c;
        }
    });
    guard.ok = true;
    ix->build = b;
    return PGVRX_OK;
}

// The insert loop shared by CREATE INDEX (disk_rule = false: the in-memory build, graph/mod.rs + build.rs) and by
// inserts into a live index (disk_rule = true: aminsert, src/index/insert.rs).  Searches, speculation and the replay
// validation are the same; what differs is how the new element's lists are chosen and how a neighbour's list is updated.
static int run_inserts(pgvrx_index *ix, const void *rows_in, int64_t n_rows, const int32_t *levels,
                       const uint64_t *heap_tids, int32_t host_threads, pgvrx_build_stats *stats, const bool disk_rule) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_rows < 0 || (n_rows > 0 && (!rows_in || !levels))) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null build input");
    if (!disk_rule && ix->loaded && !ix->build) return pgvrx_fail(ix, PGVRX_ERR_STATE, "index was loaded from arrays; build needs a handle that is empty or was built here");
    // CREATE INDEX has no notion of deleted elements; aminsert does (load_element skips them, get_update_index reuses their
    // slots in neighbour lists: insert.rs:525-528, 569-623)
    if (ix->has_flags && !disk_rule) return pgvrx_fail(ix, PGVRX_ERR_STATE, "element flags are set (vacuumed index): use pgvrx_index_insert");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    if (disk_rule && ix->loaded && !ix->build) {
        int avail0 = (int)std::max(1u, std::thread::hardware_concurrency());
        cpu_set_t set0;
        CPU_ZERO(&set0);
        if (sched_getaffinity(0, sizeof set0, &set0) == 0 && CPU_COUNT(&set0) > 0) avail0 = std::min(avail0, (int)CPU_COUNT(&set0));
        int rc0 = mirror_from_device(ix, host_threads > 0 ? host_threads : avail0);
        if (rc0) return rc0;
    }
    const int m = ix->desc.m, efc = ix->desc.ef_construction, l0 = 2 * m;
    if (!ix->build) {
        ix->build = new BuildState();
        BuildState &b = *ix->build;
        b.elem = ix->desc.elem_type; b.metric = ix->desc.metric; b.dim = ix->desc.dim; b.m = m; b.efc = efc;
        b.row_bytes = ix->row_bytes;
    }
    BuildState &g = *ix->build;
    if (g.n + n_rows > ix->desc.capacity) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "build exceeds the handle's capacity");
    const int max_level_allowed = pgvrx_max_level(m);
    for (int64_t i = 0; i < n_rows; i++)
        if (levels[i] < 0 || levels[i] > max_level_allowed) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "level out of range");

    // default: the cores this process may actually run on (a launcher may pin each rank to a subset of the machine)
    int avail = (int)std::max(1u, std::thread::hardware_concurrency());
    {
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) avail = std::min(avail, (int)CPU_COUNT(&set));
    }
    const int nthreads = host_threads > 0 ? host_threads : avail;
    Pool pool(nthreads);
    const double t_start = now_ms();
    double gpu_ms = 0.0, kernel_ms = 0.0;

    // ---- device capacity (grown once to the handle's capacity) ----
    const int64_t cap = ix->desc.capacity;
    int64_t up_needed = g.n_upper;
    for (int64_t i = 0; i < n_rows; i++) up_needed += levels[i];
    CU_TRY(ix, ix->vectors.ensure_keep(std::max<size_t>(16, (size_t)cap * ix->row_stride), (size_t)g.n * ix->row_stride));
    CU_TRY(ix, ix->adj0.ensure_keep(std::max<size_t>(16, (size_t)cap * l0 * 4), (size_t)g.n * l0 * 4));
    CU_TRY(ix, ix->adj_up.ensure_keep(std::max<size_t>(16, (size_t)up_needed * m * 4), (size_t)g.n_upper * m * 4));
    CU_TRY(ix, ix->upper_base.ensure_keep(std::max<size_t>(16, (size_t)cap * 4), (size_t)g.n * 4));
    CU_TRY(ix, ix->level.ensure_keep(std::max<size_t>(16, (size_t)cap), (size_t)g.n));
    if (ix->has_flags) {               // new elements carry no flags
        CU_TRY(ix, ix->flags.ensure_keep(std::max<size_t>(16, (size_t)cap), (size_t)g.n));
        if (cap > g.n) CU_TRY(ix, cudaMemset((uint8_t *)ix->flags.p + g.n, 0, (size_t)(cap - g.n)));
    }
    ix->generation++;                  // open iterative scans were sized for the old graph
    ix->loaded = false;
    ix->rows_loaded = false;
    ix->tc_ready = false;

    // ---- host mirror capacity ----
    // the per-element arrays are reserved for the handle's capacity once (address space only, pages are touched as
    // elements arrive), so appending calls -- pgvrx_index_insert in small batches -- never copy the mirror
    reserve_mirror(g, cap);
    auto grow_host = [&](int64_t n_new, int64_t up_new) {
        g.values.resize((size_t)n_new * g.row_bytes);
        g.level.resize((size_t)n_new);
        g.upper_base.resize((size_t)n_new);
        g.adj0.resize((size_t)n_new * l0, -1);
        g.dist0.resize((size_t)n_new * l0, 0.0f);
        g.cnt0.resize((size_t)n_new, 0);
        g.adj_up.resize((size_t)up_new * m, -1);
        g.dist_up.resize((size_t)up_new * m, 0.0f);
        g.cnt_up.resize((size_t)up_new, 0);
        g.tids.resize((size_t)n_new * PGVRX_HEAPTIDS, ~0ull);
        g.tid_cnt.resize((size_t)n_new, 0);
        g.del.resize((size_t)n_new, 0);
        g.stamp.resize((size_t)n_new, 0);
    };

    const uint8_t *rows = (const uint8_t *)rows_in;
    auto tid_of = [&](int64_t i) -> uint64_t { return heap_tids ? heap_tids[i] : (uint64_t)i; };

    // place input row i as element id e in the host mirror (no links yet)
    auto place_element = [&](int64_t i, int64_t e) {
        grow_host(e + 1, g.n_upper + levels[i]);
        memcpy(g.values.data() + (size_t)e * g.row_bytes, rows + (size_t)i * g.row_bytes, g.row_bytes);
        g.level[(size_t)e] = (uint8_t)levels[i];
        g.upper_base[(size_t)e] = levels[i] > 0 ? (int32_t)g.n_upper : -1;
        g.cnt0[(size_t)e] = 0;
        for (int j = 0; j < l0; j++) g.adj0[(size_t)e * l0 + j] = -1;
        for (int l = 0; l < levels[i]; l++) {
            g.cnt_up[(size_t)(g.n_upper + l)] = 0;
            for (int j = 0; j < m; j++) g.adj_up[(size_t)(g.n_upper + l) * m + j] = -1;
        }
        g.tid_cnt[(size_t)e] = 0;
    };

    if (!g.scratch) g.scratch = new RoundScratch();
    RoundBuffers &rb = g.scratch->rb;
    RoundScratch &rsx = *g.scratch;
    // copy `bytes` at src to page-locked staging and queue the transfer to dst on the index's stream
    auto h2d = [&](void *dst, const void *src, size_t bytes) -> int {
        if (bytes == 0) return PGVRX_OK;
        PinnedBuf &a = rsx.h_up[rsx.up_sel];
        const size_t need = rsx.up_off + ((bytes + 255) & ~(size_t)255);
        if (need > a.cap) {
            // growing frees the arena: everything queued from it must have been read first
            CU_TRY(ix, cudaStreamSynchronize(ix->stream));
            PinnedBuf bigger;
            CU_TRY(ix, bigger.ensure(std::max<size_t>(need * 2, (size_t)8 << 20)));
            if (rsx.up_off) memcpy(bigger.p, a.p, rsx.up_off);
            std::swap(a.p, bigger.p);
            std::swap(a.cap, bigger.cap);
        }
        void *stage = (uint8_t *)a.p + rsx.up_off;
        memcpy(stage, src, bytes);
        rsx.up_off = need;
        CU_TRY(ix, cudaMemcpyAsync(dst, stage, bytes, cudaMemcpyHostToDevice, ix->stream));
        return PGVRX_OK;
    };
    auto next_arena = [&]() { rsx.up_sel ^= 1; rsx.up_off = 0; };
    CU_TRY(ix, cudaStreamSynchronize(ix->stream));       // nothing an earlier call queued still reads the arenas
    rsx.up_off = 0;

    // scatter staging on the host
    std::vector<int64_t> sc_off0, sc_offu;
    std::vector<int32_t> sc_rows0, sc_rowsu;
    auto queue_row = [&](int64_t v, int lc) {
        if (lc == 0) {
            sc_off0.push_back(v * l0);
            sc_rows0.insert(sc_rows0.end(), &g.adj0[(size_t)v * l0], &g.adj0[(size_t)v * l0] + l0);
        } else {
            int64_t r = g.upper_base[(size_t)v] + lc - 1;
            sc_offu.push_back(r * m);
            sc_rowsu.insert(sc_rowsu.end(), &g.adj_up[(size_t)r * m], &g.adj_up[(size_t)r * m] + m);
        }
    };
    auto flush_scatter = [&]() -> int {
        for (int pass = 0; pass < 2; pass++) {
            std::vector<int64_t> &off = pass == 0 ? sc_off0 : sc_offu;
            std::vector<int32_t> &dat = pass == 0 ? sc_rows0 : sc_rowsu;
            if (off.empty()) continue;
            const int row_len = pass == 0 ? l0 : m;
            CU_TRY(ix, rb.sc_off.ensure(off.size() * 8));
            CU_TRY(ix, rb.sc_rows.ensure(dat.size() * 4));
            // the two passes share rb.sc_off / rb.sc_rows: the second pass's copies queue behind the first pass's kernel
            { int rc_ = h2d(rb.sc_off.p, off.data(), off.size() * 8); if (rc_) return rc_; }
            { int rc_ = h2d(rb.sc_rows.p, dat.data(), dat.size() * 4); if (rc_) return rc_; }
            CU_TRY(ix, launch_scatter_rows((int32_t *)(pass == 0 ? ix->adj0.p : ix->adj_up.p), (const int64_t *)rb.sc_off.p,
                                           (const int32_t *)rb.sc_rows.p, row_len, (int)off.size(), ix->stream));
            off.clear();
            dat.clear();
        }
        return PGVRX_OK;
    };
    // upload vector rows / level / upper_base of elements [e0, e1)
    auto upload_elements = [&](int64_t e0, int64_t e1) -> int {
        if (e1 <= e0) return PGVRX_OK;
        if (ix->row_stride == ix->row_bytes) {
            int rc_ = h2d((uint8_t *)ix->vectors.p + (size_t)e0 * ix->row_stride, g.row(e0), (size_t)(e1 - e0) * g.row_bytes);
            if (rc_) return rc_;
        } else {
            CU_TRY(ix, cudaMemsetAsync((uint8_t *)ix->vectors.p + (size_t)e0 * ix->row_stride, 0, (size_t)(e1 - e0) * ix->row_stride, ix->stream));
            CU_TRY(ix, cudaMemcpy2DAsync((uint8_t *)ix->vectors.p + (size_t)e0 * ix->row_stride, ix->row_stride, g.row(e0), g.row_bytes,
                                         g.row_bytes, (size_t)(e1 - e0), cudaMemcpyHostToDevice, ix->stream));
        }
        { int rc_ = h2d((uint8_t *)ix->level.p + e0, &g.level[(size_t)e0], (size_t)(e1 - e0)); if (rc_) return rc_; }
        { int rc_ = h2d((int32_t *)ix->upper_base.p + e0, &g.upper_base[(size_t)e0], (size_t)(e1 - e0) * 4); if (rc_) return rc_; }
        return PGVRX_OK;
    };

    int64_t f = 0;                       // next input row
    uint64_t n_rounds = 0, n_searched = 0, n_conflicts = 0, n_dups = 0, n_retry_big = 0;
    uint64_t c_dist = 0, c_exp0 = 0, c_expu = 0;

    // the first element of an empty index becomes the entry point without a search (build.rs:526-529)
    // (so does the first element inserted after a vacuum that left no entry point behind: insert.rs:1317-1334 finds no
    // neighbours without one, :1454-1470 makes the element the entry point)
    if (g.entry < 0 && n_rows > 0) {
        const int64_t e0 = g.n;
        place_element(0, e0);
        g.tids[(size_t)e0 * PGVRX_HEAPTIDS] = tid_of(0);
        g.tid_cnt[(size_t)e0] = 1;
        g.n_upper += levels[0];
        g.n = e0 + 1;
        g.entry = e0;
        g.entry_level = levels[0];
        int rc = upload_elements(e0, e0 + 1);
        if (rc) return rc;
        queue_row(e0, 0);
        for (int l = 1; l <= levels[0]; l++) queue_row(e0, l);
        rc = flush_scatter();
        if (rc) return rc;
        f = 1;
    }

    int window = 4;
    const int WINDOW_MAX = 512;
    // CTA-per-search kernel: on for wide rows (an expansion is several gather rounds: 20k x 768-d halfvec rows, kernels
    // 27.7 s -> 18.6 s), off for rows that are staged whole -- there nearly every popped candidate is one the previous
    // expansion has just inserted, nothing can be computed ahead of it, and the two kernels take the same time (C2 100k
    // rows: 5.84 s vs 5.93 s).  PGVRX_BUILD_TEAM=0/1 pins the choice.
    bool team_on = ix->row_stride > 1056 && !ix->has_flags;     // (the CTA-per-search kernel has no flagged form)
    if (const char *e = getenv("PGVRX_BUILD_TEAM")) team_on = atoi(e) != 0;
    uint64_t n_team_rounds = 0;
    // scratch escalation for a search that overflowed its visited table / tie tail: the head of the round is re-run alone
    // with 16x, 256x, 4096x the visited table (the ef = 1 descent then uses the whole table too) and 2048 / 8192 / 20000
    // tie slots; if even that is not enough the call fails instead of repeating the same round for ever
    int big_level = 0;
    const int BIG_MAX = 3;
    int big_tail[4] = {-1, 2048, 8192, 20000};
    if (const char *e = getenv("PGVRX_BUILD_TAILS")) {      // "a,b,c": tie slots of the three escalation levels (tests)
        int a = 0, b = 0, c = 0;
        if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a >= 1 && b >= a && c >= b && c <= 20000) { big_tail[1] = a; big_tail[2] = b; big_tail[3] = c; }
    }
    const int rs_cap = std::max(256, 6 * efc + 64);
    const int rs_ent = 3 + (l0 + 31) / 32;               // node, f_dist bits, layer, near-mask words of the list read
    struct GroupPart {                                  // groups of the nodes with node % NP == part; reused across rounds
        std::vector<Group> groups;
        int n = 0;
        std::unordered_map<uint64_t, int> up;           // (node, layer >= 1) -> group; layer 0 uses the stamp arrays
    };
    const uint32_t NP = 16;
    std::vector<GroupPart> parts(NP);
    std::vector<Group *> all_groups;
    std::vector<uint32_t> gstamp0;
    std::vector<int32_t> gidx0;
    std::vector<int32_t> h_gids, h_goff;
    std::vector<float> h_gown;
    PinnedBuf &h_mat = g.scratch->h_mat;
    uint64_t n_pair_gpu = 0, n_val_dist = 0;
    std::atomic<uint64_t> n_val_dist_atomic{0};
    double t_sim = 0, t_g1 = 0, t_g2 = 0;
    for (int t = 0; t < 256; t++) g_identity_pos[t] = (int16_t)t;
    uint64_t why_overflow = 0, why_notfull = 0, why_pending = 0, why_order = 0;   // conflict reasons (PGVRX_BUILD_DEBUG)
    double t_phase_a = 0, t_groups = 0, t_commit = 0, t_scatter = 0, t_prelude = 0, t_tail = 0;
    const double t_loop0 = now_ms();
    double t_place = 0, t_upload = 0, t_cfg = 0;
    // rows a pairwise group may hold: what fits in one CTA's shared memory, at most PAIR_MAX_ROWS
    const uint32_t pair_stride = ix->row_stride + (((ix->row_stride / 16) & 1u) ? 0u : 16u);
    const int pair_rows = (int)std::min<size_t>(PAIR_MAX_ROWS, (200 * 1024) / pair_stride);
    const int pair_occ = std::max(1, (int)((220 * 1024) / (128 + (size_t)pair_rows * pair_stride)));
    const bool w_masks = !disk_rule && efc <= pair_rows;  // select_neighbors of the new elements answered by GPU masks
    std::vector<int32_t> h_elem;
    PinnedBuf &h_pack = g.scratch->h_pack;

    while (f < n_rows) {
        const int B = (int)std::min<int64_t>(big_level > 0 ? 1 : window, n_rows - f);
        const double tP0 = now_ms();
        g.round++;
        n_rounds++;
        next_arena();
        // ---- provisional placement: input row f+i -> element id g.n + i (valid while no duplicate commits) ----
        const int64_t e_base = g.n;
        const int64_t up_base = g.n_upper;
        int max_layers = 1;
        {
            int64_t up = up_base;
            for (int i = 0; i < B; i++) {
                const int64_t save = g.n_upper;
                g.n_upper = up;                       // place_element reads g.n_upper as this element's first upper row
                place_element(f + i, e_base + i);
                g.n_upper = save;
                up += levels[f + i];
                max_layers = std::max(max_layers, std::min(levels[f + i], g.entry_level) + 1);
            }
        }
        const double tP1 = now_ms();
        t_place += tP1 - tP0;
        int rc = upload_elements(e_base, e_base + B);
        if (rc) return rc;
        const double tP2 = now_ms();
        t_upload += tP2 - tP1;

        // ---- K2 on the GPU ----
        ScanConfig cfg;
        uint32_t vis_entries = pgvrx_pow2_at_least(std::max<uint64_t>(1024, (uint64_t)efc * 2 * m * 4));
        if (big_level > 0)
            vis_entries = pgvrx_pow2_at_least(std::min<uint64_t>((uint64_t)vis_entries << (4 * big_level), std::max<uint64_t>(1024, (uint64_t)cap * 2)));
        rc = pgvrx_make_config(ix, efc, big_tail[big_level], vis_entries, B, 0, true, &cfg);
        if (rc) return rc;
        const int wpb = cfg.wpb;
        CU_TRY(ix, rb.vis.ensure((size_t)cfg.grid * wpb * cfg.vis_entries * 4));
        CU_TRY(ix, rb.elem_ids.ensure((size_t)B * 4));
        // every per-round result lives in one device block mirrored by one pinned host block: a single D2H copy per round
        size_t po = 0;
        auto carve = [&](size_t bytes) { const size_t o = po; po += (bytes + 255) & ~(size_t)255; return o; };
        const size_t o_wids = carve((size_t)B * max_layers * efc * 4), o_wd = carve((size_t)B * max_layers * efc * 4);
        const size_t o_wcnt = carve((size_t)B * max_layers * 4), o_rs = carve((size_t)B * rs_cap * rs_ent * 4);
        const size_t o_rscnt = carve((size_t)B * 4), o_status = carve((size_t)B * 4), o_ctr = carve((size_t)B * 16);
        const size_t o_wmask = carve(w_masks ? (size_t)B * max_layers * efc * (PAIR_MAX_ROWS / 32) * 4 : 0);
        const size_t o_pp = carve((size_t)B * B * 4);          // d(pending j, pending i), i < j
        CU_TRY(ix, rb.out_pack.ensure(po));
        CU_TRY(ix, h_pack.ensure(po));
        uint8_t *dpk = (uint8_t *)rb.out_pack.p;
        const uint8_t *hpk = (const uint8_t *)h_pack.p;
        CU_TRY(ix, rb.work.ensure(64));
        t_cfg += now_ms() - tP2;
        h_elem.resize((size_t)B);
        for (int i = 0; i < B; i++) h_elem[(size_t)i] = (int32_t)(e_base + i);
        { int rc_ = h2d(rb.elem_ids.p, h_elem.data(), (size_t)B * 4); if (rc_) return rc_; }
        CU_TRY(ix, cudaMemsetAsync(rb.work.p, 0, 64, ix->stream));
        ix->n = e_base;                 // snapshot: only committed elements are reachable, but rows of the pending ones are resident
        ix->entry = g.entry;
        ix->entry_level = g.entry_level;
        GraphView gv = pgvrx_view_of(ix);
        BuildParams bp;
        bp.elem_ids = (const int32_t *)rb.elem_ids.p;
        bp.n_pending = B;
        bp.efc = efc;
        bp.tail_cap = cfg.L.tail_cap;
        bp.max_layers = max_layers;
        bp.w_ids = (int32_t *)(dpk + o_wids);
        bp.w_dists = (float *)(dpk + o_wd);
        bp.w_counts = (int32_t *)(dpk + o_wcnt);
        bp.rs_log = (int32_t *)(dpk + o_rs);
        bp.rs_cap = rs_cap;
        bp.rs_ent = rs_ent;
        bp.rs_counts = (int32_t *)(dpk + o_rscnt);
        bp.out_status = (uint32_t *)(dpk + o_status);
        bp.out_ctr = (uint32_t *)(dpk + o_ctr);
        bp.vis_base = (uint32_t *)rb.vis.p;
        bp.vis_entries = cfg.vis_entries;
        bp.up_entries = big_level > 0 ? cfg.vis_entries : std::min<uint32_t>(cfg.vis_entries, 1024u);
        bp.work_counter = (int *)rb.work.p;
        bp.stage_stride = cfg.L.stage_stride;
        bp.R = cfg.L.R;
        bp.stage_bytes = cfg.L.stage_bytes;
        bp.chunk_bytes = cfg.L.chunk_bytes;
        bp.chunk_stride = cfg.L.chunk_stride;
        const double tg0 = now_ms();
        t_prelude += tg0 - tP0;
        CU_TRY(ix, cudaEventRecord(ix->ev0, ix->stream));
        // the first attempt of every search runs on the CTA-per-search kernel (one master warp + helper warps, visited
        // set in shared memory: single-search latency is what bounds a round); escalation levels use the one-warp kernel
        TeamGeom tg;
        size_t team_smem = 0;
        int team_grid = 0;
        const bool use_team = team_on && !ix->has_flags && big_level == 0 && team_config(ix, cfg, efc, B, &tg, &team_smem, &team_grid);
        if (use_team) {
            CU_TRY(ix, launch_team_build(ix->desc.elem_type, ix->desc.metric, cfg.L.list_j, gv, bp, tg, team_grid, team_smem, ix->stream));
            n_team_rounds++;
        } else {
            CU_TRY(ix, launch_build_search(ix->desc.elem_type, ix->desc.metric, cfg.L.list_j, gv, bp, cfg.grid, cfg.smem_block, cfg.wpb * 32, ix->stream));
        }
        if (w_masks) {
            // check_element_closer for the candidate lists the searches just produced (graph/mod.rs:419 -> :315-339)
            CU_TRY(ix, cudaMemsetAsync((int *)rb.work.p + 4, 0, 4, ix->stream));
            PairParams wp;
            wp.ids = bp.w_ids;
            wp.own_dist = bp.w_dists;
            wp.group_off = nullptr;
            wp.w_counts = bp.w_counts;
            wp.w_pitch = efc;
            wp.w_layers = max_layers;
            wp.n_groups = B * max_layers;
            wp.out = (uint32_t *)(dpk + o_wmask);
            wp.work_counter = (int *)rb.work.p + 4;
            wp.stage_stride = cfg.L.stage_stride;
            CU_TRY(ix, launch_pairwise(ix->desc.elem_type, ix->desc.metric, gv, wp, pair_rows, std::min(B * max_layers, ix->sm_count * pair_occ), ix->stream));
        }
        CU_TRY(ix, launch_pending_pairs(ix->desc.elem_type, ix->desc.metric, gv, bp.elem_ids, B, (float *)(dpk + o_pp), ix->sm_count, ix->stream));
        CU_TRY(ix, cudaEventRecord(ix->ev1, ix->stream));
        CU_TRY(ix, cudaMemcpyAsync(h_pack.p, rb.out_pack.p, po, cudaMemcpyDeviceToHost, ix->stream));
        const int32_t *h_wids = (const int32_t *)(hpk + o_wids);
        const float *h_wd = (const float *)(hpk + o_wd);
        const int32_t *h_wcnt = (const int32_t *)(hpk + o_wcnt);
        const int32_t *h_rs = (const int32_t *)(hpk + o_rs);
        const int32_t *h_rscnt = (const int32_t *)(hpk + o_rscnt);
        const uint32_t *h_status = (const uint32_t *)(hpk + o_status);
        const uint32_t *h_ctr = (const uint32_t *)(hpk + o_ctr);
        const uint32_t *h_wmask = (const uint32_t *)(hpk + o_wmask);
        const float *h_pp = (const float *)(hpk + o_pp);
        CU_TRY(ix, cudaStreamSynchronize(ix->stream));
        gpu_ms += now_ms() - tg0;
        {
            float kms = 0.0f;
            CU_TRY(ix, cudaEventElapsedTime(&kms, ix->ev0, ix->ev1));
            kernel_ms += kms;
        }
        n_searched += (uint64_t)B;

        // ---- host: phase A (parallel): select_neighbors + duplicate probe for every searched element ----
        const double tA0 = now_ms();
        std::vector<Pending> pend((size_t)B);
        pool.parallel_for(B, [&](int64_t i) {
            Pending &p = pend[(size_t)i];
            p.input_row = f + i;
            p.level = levels[f + i];
            if (h_status[(size_t)i] != 0) return;
            const int top = std::min(p.level, g.entry_level);
            p.sel.resize((size_t)p.level + 1);
            std::vector<Cand> w;
            std::vector<int16_t> wpos, spos;
            for (int lc = top; lc >= 0; lc--) {
                const int cnt = h_wcnt[(size_t)i * max_layers + lc];
                w.resize((size_t)std::max(cnt, 0));
                const size_t o = ((size_t)i * max_layers + lc) * efc;
                for (int t = 0; t < cnt; t++) w[(size_t)t] = Cand{h_wd[o + t], h_wids[o + t]};
                if (disk_rule) {                                                     // insert.rs:1110-1116: the nearest lm, no heuristic
                    p.sel[(size_t)lc].assign(w.begin(), w.begin() + std::min(cnt < 0 ? 0 : cnt, g.lm(lc)));
                } else if (w_masks && cnt > g.lm(lc)) {                                     // graph/mod.rs:419 with the GPU's masks
                    wpos.resize((size_t)cnt);
                    for (int t = 0; t < cnt; t++) wpos[(size_t)t] = (int16_t)t;
                    select_neighbors_mask(g, w.data(), wpos.data(), cnt, g.lm(lc), h_wmask + o * (PAIR_MAX_ROWS / 32), p.sel[(size_t)lc], spos);
                } else {
                    select_neighbors(g, w.data(), cnt, g.lm(lc), p.sel[(size_t)lc]);
                }
            }
            // duplicate probe (build.rs:482-503): walk the selected layer-0 neighbours while distance == 0
            for (const Cand &nb : p.sel[0]) {
                if (nb.distance != 0.0f) break;
                if (memcmp(g.row(e_base + i), g.row(nb.idx), g.row_bytes) == 0) { p.dup_of = nb.idx; break; }
            }
        });

        // ---- back-link groups: every (node, layer) some pending element selected, with the pending elements
        //      that selected it in insert order.  Where a full list will have to be re-selected
        //      (graph/mod.rs:473-485) the pairwise distances among list + newcomers come from the GPU. ----
        const double tG0 = now_ms();
        t_phase_a += tG0 - tA0;
        // groups are built by NP workers, each owning the nodes with node % NP == its part (no shared writes)
        int n_ok = 0;
        while (n_ok < B && h_status[(size_t)n_ok] == 0) n_ok++;      // nothing after an overflowed search can commit this round
        if ((int64_t)gstamp0.size() < cap) { gstamp0.resize((size_t)cap, 0); gidx0.resize((size_t)cap, -1); }
        auto find_group = [&](int32_t node, int lc) -> Group * {
            GroupPart &P = parts[(size_t)((uint32_t)node % NP)];
            if (lc == 0) return gstamp0[(size_t)node] == g.round ? &P.groups[(size_t)gidx0[(size_t)node]] : nullptr;
            auto it = P.up.find(((uint64_t)(uint32_t)node << 8) | (uint64_t)lc);
            return it == P.up.end() ? nullptr : &P.groups[(size_t)it->second];
        };
        pool.parallel_for(NP, [&](int64_t part) {
            GroupPart &P = parts[(size_t)part];
            P.n = 0;
            P.up.clear();
            for (int i = 0; i < n_ok; i++) {
                const Pending &p = pend[(size_t)i];
                for (int lc = 0; lc < (int)p.sel.size(); lc++)
                    for (const Cand &nb : p.sel[(size_t)lc]) {
                        if ((int64_t)((uint32_t)nb.idx % NP) != part) continue;
                        int gi = -1;
                        if (lc == 0) { if (gstamp0[(size_t)nb.idx] == g.round) gi = gidx0[(size_t)nb.idx]; }
                        else { auto it = P.up.find(((uint64_t)(uint32_t)nb.idx << 8) | (uint64_t)lc); if (it != P.up.end()) gi = it->second; }
                        if (gi < 0) {
                            gi = P.n++;
                            if ((int)P.groups.size() < P.n) P.groups.emplace_back();
                            if (lc == 0) { gstamp0[(size_t)nb.idx] = g.round; gidx0[(size_t)nb.idx] = gi; }
                            else P.up.emplace(((uint64_t)(uint32_t)nb.idx << 8) | (uint64_t)lc, gi);
                            Group &G = P.groups[(size_t)gi];
                            G.node = nb.idx;
                            G.lc = lc;
                            G.n_list = g.cnt(nb.idx, lc);
                            const int32_t *ids = g.ids(nb.idx, lc);
                            const float *ds = g.dists(nb.idx, lc);
                            G.ids.assign(ids, ids + G.n_list);
                            G.own.assign(ds, ds + G.n_list);
                            G.lm = g.lm(lc);
                            G.step = 0;
                            G.modified = false;
                        }
                        P.groups[(size_t)gi].ids.push_back((int32_t)(e_base + i));
                        P.groups[(size_t)gi].own.push_back(nb.distance);
                    }
            }
        });
        t_g1 += now_ms() - tG0;
        const double tG1 = now_ms();
        all_groups.clear();
        h_goff.clear();
        size_t n_rows_dev = 0;
        int n_dev_groups = 0;
        for (GroupPart &P : parts)
            for (int gi = 0; gi < P.n; gi++) {
                Group &G = P.groups[(size_t)gi];
                all_groups.push_back(&G);
                const int total = (int)G.ids.size();
                G.row_off = -1;
                if (total > G.lm && total <= pair_rows) {           // a re-selection can happen and the group fits on a CTA
                    G.row_off = (int)n_rows_dev;
                    h_goff.push_back((int32_t)n_rows_dev);
                    n_rows_dev += (size_t)total;
                    n_dev_groups++;
                }
            }
        const int n_groups = (int)all_groups.size();
        t_g2 += now_ms() - tG1;
        h_gids.resize(n_rows_dev);
        h_gown.resize(n_rows_dev);
        pool.parallel_for(n_groups, [&](int64_t gi) {
            const Group &G = *all_groups[(size_t)gi];
            if (G.row_off < 0) return;
            std::copy(G.ids.begin(), G.ids.end(), h_gids.begin() + G.row_off);
            std::copy(G.own.begin(), G.own.end(), h_gown.begin() + G.row_off);
        });
        h_goff.push_back((int32_t)n_rows_dev);
        t_groups += now_ms() - tG0;
        constexpr int MW = PAIR_MAX_ROWS / 32;
        if (n_dev_groups > 0) {
            const double tp0 = now_ms();
            CU_TRY(ix, rb.g_ids.ensure(n_rows_dev * 4));
            CU_TRY(ix, rb.g_own.ensure(n_rows_dev * 4));
            CU_TRY(ix, rb.g_off.ensure(h_goff.size() * 4));
            CU_TRY(ix, rb.g_mat.ensure(n_rows_dev * MW * 4));
            CU_TRY(ix, h_mat.ensure(n_rows_dev * MW * 4));
            { int rc_ = h2d(rb.g_ids.p, h_gids.data(), n_rows_dev * 4); if (rc_) return rc_; }
            { int rc_ = h2d(rb.g_own.p, h_gown.data(), n_rows_dev * 4); if (rc_) return rc_; }
            { int rc_ = h2d(rb.g_off.p, h_goff.data(), h_goff.size() * 4); if (rc_) return rc_; }
            CU_TRY(ix, cudaMemsetAsync((int *)rb.work.p + 8, 0, 4, ix->stream));
            PairParams pp;
            pp.ids = (const int32_t *)rb.g_ids.p;
            pp.own_dist = (const float *)rb.g_own.p;
            pp.group_off = (const int32_t *)rb.g_off.p;
            pp.w_counts = nullptr;
            pp.w_pitch = pp.w_layers = 0;
            pp.n_groups = n_dev_groups;
            pp.out = (uint32_t *)rb.g_mat.p;
            pp.work_counter = (int *)rb.work.p + 8;
            pp.stage_stride = cfg.L.stage_stride;
            CU_TRY(ix, cudaEventRecord(ix->ev0, ix->stream));
            CU_TRY(ix, launch_pairwise(ix->desc.elem_type, ix->desc.metric, gv, pp, pair_rows, std::min(n_dev_groups, ix->sm_count * pair_occ), ix->stream));
            CU_TRY(ix, cudaEventRecord(ix->ev1, ix->stream));
            CU_TRY(ix, cudaMemcpyAsync(h_mat.p, rb.g_mat.p, n_rows_dev * MW * 4, cudaMemcpyDeviceToHost, ix->stream));
            gpu_ms += now_ms() - tp0;
        }

        if (n_dev_groups > 0) {
            const double tp1 = now_ms();
            CU_TRY(ix, cudaStreamSynchronize(ix->stream));
            float kms = 0.0f;
            CU_TRY(ix, cudaEventElapsedTime(&kms, ix->ev0, ix->ev1));
            kernel_ms += kms;
            gpu_ms += now_ms() - tp1;
            for (const Group *G : all_groups)
                if (G->row_off >= 0) n_pair_gpu += (uint64_t)G->ids.size() * G->ids.size();
        }
        const uint32_t *mask_all = (const uint32_t *)h_mat.p;

        // ---- host (parallel by group): update_neighbor_connections for every group, all members in order ----
        const double tB0 = now_ms();
        pool.parallel_for(n_groups, [&](int64_t gi) {
            Group &G = *all_groups[(size_t)gi];
            if (disk_rule) simulate_group_disk(g, G, G.row_off >= 0 ? mask_all + (size_t)G.row_off * MW : nullptr, ix->has_flags, e_base);
            else simulate_group(g, G, G.row_off >= 0 ? mask_all + (size_t)G.row_off * MW : nullptr);
        });
        t_sim += now_ms() - tB0;

        // ---- host: sequential validation + commit in insert order ----
        const double tC0 = now_ms();
        int committed = 0;
        bool stop = false;
        for (int i = 0; i < B && !stop; i++) {
            Pending &p = pend[(size_t)i];
            const int64_t e = e_base + i;
            if (h_status[(size_t)i] != 0) {                     // scratch overflow: redo alone with large scratch
                if (i == 0) {
                    if (big_level >= BIG_MAX) {
                        char b[200];
                        snprintf(b, sizeof b, "the search for input row %lld exhausted the largest search scratch (status %u: %s)",
                                 (long long)(f + i), h_status[(size_t)i],
                                 (h_status[(size_t)i] & 2u) ? "too many candidates tie with the furthest result" : "visited table");
                        grow_host(g.n, g.n_upper);          // drop the provisional elements of this round
                        ix->n = g.n; ix->n_upper = g.n_upper; ix->entry = g.entry; ix->entry_level = g.entry_level;
                        ix->loaded = g.n > 0; ix->rows_loaded = g.n > 0;
                        return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, b);
                    }
                    big_level++;
                    n_retry_big++;
                }
                break;
            }
            // -- is the speculative search still the one a sequential build would have run? --
            // Replay argument: the search is a function of the lists it read.  For a list that changed since the
            // snapshot, compare what it could contribute at the moment it was read: only neighbours nearer than
            // W's furthest distance at that expansion (f_t, W full) can be accepted then or ever after (f only
            // shrinks; equal is rejected, graph/mod.rs:226), everything else merely toggles a visited flag of an
            // element that is rejected wherever it is met.  If the acceptable neighbours of the old and the new
            // list are the same sequence, the run is identical.
            bool valid = true;
            if (i > 0) {
                if (h_rscnt[(size_t)i] > rs_cap) { valid = false; why_overflow++; }
                const int32_t *rs = &h_rs[(size_t)i * rs_cap * rs_ent];
                for (int t = 0; valid && t < h_rscnt[(size_t)i]; t++) {
                    const int32_t *ent = rs + (size_t)rs_ent * t;
                    const int32_t v = ent[0], lc = ent[2];
                    if (g.stamp[(size_t)v] != g.round) continue;
                    const Group *Gp = find_group(v, lc);
                    if (!Gp) continue;
                    const Group &G = *Gp;
                    if (G.step == 0) continue;                  // nobody has linked to it yet
                    const int n_old = G.n_list;
                    const uint32_t ford = (uint32_t)ent[1];
                    const int32_t *l_new;
                    const float *ds_new;
                    const int16_t *p_new;
                    int n_new;
                    group_list(G, l_new, ds_new, p_new, n_new);
                    if (ford == 0xFFFFFFFFu) {                  // W not full: any change of the list matters
                        valid = n_new == n_old && std::equal(l_new, l_new + n_new, G.ids.data());
                        if (!valid) why_notfull++;
                        continue;
                    }
                    // d(e, member) < f_t, exactly as the kernel compares: for the entries of the list the search read the
                    // kernel logged the answer (it evaluated every entry of the row); newcomers are evaluated here
                    const uint32_t *bits = reinterpret_cast<const uint32_t *>(ent + 3);
                    float f_t;
                    { const uint32_t fb = (ford & 0x80000000u) ? (ford & 0x7FFFFFFFu) : ~ford; memcpy(&f_t, &fb, 4); }
                    auto near = [&](int pos) -> bool {
                        if (pos < n_old) return (bits[pos >> 5] >> (pos & 31)) & 1u;
                        // a newcomer is a pending element of this round: its distance to e came with the round's results
                        n_val_dist++;
                        return ord_of(h_pp[(size_t)i * B + (size_t)(G.ids[(size_t)pos] - (int32_t)e_base)]) < ford;
                    };
                    int io = 0, in_ = 0;
                    while (valid) {
                        while (io < n_old && !near(io)) io++;
                        while (in_ < n_new && !near(p_new[in_])) in_++;
                        const bool eo = io >= n_old, en = in_ >= n_new;
                        if (eo || en) {
                            valid = eo && en;
                            if (!valid) { if (!en && l_new[in_] >= e_base) why_pending++; else why_order++; }
                            break;
                        }
                        if (io != p_new[in_]) { valid = false; if (l_new[in_] >= e_base) why_pending++; else why_order++; }
                        io++;
                        in_++;
                    }
                }
            }
            if (!valid) { n_conflicts++; break; }
            c_dist += h_ctr[(size_t)i * 4];
            c_exp0 += h_ctr[(size_t)i * 4 + 1];
            c_expu += h_ctr[(size_t)i * 4 + 2];
            big_level = 0;
            committed++;
            // duplicate merge (build.rs:495-512): the probe skipped the TID-capacity test, apply it in order here
            bool merged = false;
            if (p.dup_of >= 0) {
                for (const Cand &nb : p.sel[0]) {
                    if (nb.distance != 0.0f) break;
                    // insert.rs:1159-1163 also refuses an element whose first heap TID is gone (being deleted)
                    if (memcmp(g.row(e), g.row(nb.idx), g.row_bytes) == 0 && g.tid_cnt[(size_t)nb.idx] < PGVRX_HEAPTIDS &&
                        !(disk_rule && g.tid_cnt[(size_t)nb.idx] == 0)) {
                        g.tids[(size_t)nb.idx * PGVRX_HEAPTIDS + g.tid_cnt[(size_t)nb.idx]++] = tid_of(p.input_row);
                        merged = true;
                        break;
                    }
                }
            }
            if (merged) {                                       // element is dropped: later provisional ids shift
                n_dups++;
                stop = true;
                break;
            }
            // new element's own lists (graph/mod.rs:422), then its back-links at once (graph/mod.rs:442-489):
            // later elements of this round are validated against the lists as they stand after this commit
            const int top = std::min(p.level, g.entry_level);
            for (int lc = top; lc >= 0; lc--) {
                const std::vector<Cand> &s = p.sel[(size_t)lc];
                int32_t *ids = g.ids(e, lc);
                float *ds = g.dists(e, lc);
                for (size_t t = 0; t < s.size(); t++) { ids[t] = s[t].idx; ds[t] = s[t].distance; }
                g.cnt(e, lc) = (uint8_t)s.size();
            }
            for (int lc = std::min(p.level, top); lc >= 0; lc--) {   // :455 for lc in (0..=level).rev(); lists above `top` are empty
                for (const Cand &nb : p.sel[(size_t)lc]) {
                    Group &G = *find_group(nb.idx, lc);
                    // commits are a prefix of the batch and a group's members are in batch order: this is the next one
                    G.step++;
                    G.modified = true;
                    g.stamp[(size_t)nb.idx] = g.round;
                }
            }
            g.tids[(size_t)e * PGVRX_HEAPTIDS] = tid_of(p.input_row);
            g.tid_cnt[(size_t)e] = 1;
            g.n = e + 1;
            g.n_upper = g.upper_base[(size_t)e] >= 0 ? g.upper_base[(size_t)e] + p.level : g.n_upper;
            if (p.level > g.entry_level) {                      // build.rs:523-525: strictly greater
                g.entry = e;
                g.entry_level = p.level;
                stop = true;                                    // later searches started from the old entry point
            }
        }

        // ---- device graph := committed state ----
        const double tS0 = now_ms();
        t_commit += tS0 - tC0;
        for (int gi = 0; gi < n_groups; gi++) {
            const Group &G = *all_groups[(size_t)gi];
            if (!G.modified) continue;
            const int32_t *l;
            const float *d;
            const int16_t *pp_;
            int cnt;
            group_list(G, l, d, pp_, cnt);
            int32_t *ids = g.ids(G.node, G.lc);
            float *ds = g.dists(G.node, G.lc);
            for (int t = 0; t < cnt; t++) { ids[t] = l[t]; ds[t] = d[t]; }
            for (int t = cnt; t < G.lm; t++) ids[t] = -1;
            g.cnt(G.node, G.lc) = (uint8_t)cnt;
            queue_row(G.node, G.lc);
        }
        for (int64_t e = e_base; e < g.n; e++) {
            queue_row(e, 0);
            for (int l = 1; l <= g.level[(size_t)e]; l++) queue_row(e, l);
        }
        rc = flush_scatter();
        if (rc) return rc;
        t_scatter += now_ms() - tS0;

        // input rows consumed this round: committed elements plus a merged duplicate
        const double tT0 = now_ms();
        f += committed;
        if (big_level == 0) window = std::min(WINDOW_MAX, std::max(4, 2 * committed + 2));
        // host arrays beyond g.n hold provisional (uncommitted) elements; shrink the logical size back
        grow_host(g.n, g.n_upper);
        t_tail += now_ms() - tT0;
    }

    if (getenv("PGVRX_BUILD_DEBUG"))
        fprintf(stderr, "[pgvrx build] rounds on the CTA-per-search kernel: %llu of %llu\n", (unsigned long long)n_team_rounds, (unsigned long long)n_rounds);
    if (getenv("PGVRX_BUILD_DEBUG"))
        fprintf(stderr, "[pgvrx build] conflicts: log overflow %llu, W-not-full %llu, pending element acceptable %llu, order/removal %llu | ms: gpu %.0f (kernels %.0f) phaseA %.0f groups %.0f (build %.0f enum %.0f) simulate %.0f commit %.0f scatter %.0f | setup %.0f prelude %.0f (place %.0f upload %.0f config+buffers %.0f) tail %.0f total %.0f\n",
                (unsigned long long)why_overflow, (unsigned long long)why_notfull, (unsigned long long)why_pending,
                (unsigned long long)why_order, gpu_ms, kernel_ms, t_phase_a, t_groups, t_g1, t_g2, t_sim, t_commit, t_scatter, t_loop0 - t_start, t_prelude, t_place, t_upload, t_cfg, t_tail,
                now_ms() - t_start);
    ix->n = g.n;
    ix->n_upper = g.n_upper;
    ix->entry = g.entry;
    ix->entry_level = g.entry_level;
    ix->loaded = true;
    ix->rows_loaded = true;
    if (stats) {
        stats->n_rounds = n_rounds;
        stats->n_searched = n_searched;
        stats->n_conflicts = n_conflicts;
        stats->n_duplicates = n_dups;
        stats->n_big_retries = n_retry_big;
        stats->n_dist = c_dist;
        stats->n_expand0 = c_exp0;
        stats->n_expand_upper = c_expu;
        stats->n_pair = g.n_pair.load();
        stats->n_pair_gpu = n_pair_gpu;
        stats->n_host_dist = g.n_host_dist.load();
        stats->n_validation_dist = n_val_dist + n_val_dist_atomic.load();
        stats->gpu_ms = gpu_ms;
        stats->kernel_ms = kernel_ms;
        stats->total_ms = now_ms() - t_start;
        stats->host_threads = nthreads;
    }
    return PGVRX_OK;
}

// the host mirror could not grow: the graph may be half-updated, the handle is only good for pgvrx_index_destroy
static int host_oom(pgvrx_index *ix) {
    return ix ? pgvrx_fail(ix, PGVRX_ERR_NOMEM, "out of host memory while updating the index (destroy the handle)") : PGVRX_ERR_NOMEM;
}

int pgvrx_index_build(pgvrx_index *ix, const void *rows, int64_t n_rows, const int32_t *levels, const uint64_t *heap_tids,
                      int32_t host_threads, pgvrx_build_stats *stats) {
    try {
        return run_inserts(ix, rows, n_rows, levels, heap_tids, host_threads, stats, false);
    } catch (const std::bad_alloc &) {                   // no exception crosses the C ABI
        return host_oom(ix);
    }
}

int pgvrx_index_insert(pgvrx_index *ix, const void *rows, int64_t n_rows, const int32_t *levels, const uint64_t *heap_tids,
                       int32_t host_threads, pgvrx_build_stats *stats) {
    try {
        return run_inserts(ix, rows, n_rows, levels, heap_tids, host_threads, stats, true);
    } catch (const std::bad_alloc &) {
        return host_oom(ix);
    }
}

// ------------------------------------------------------------------------------------------------
// VACUUM: ambulkdelete (src/index/vacuum.rs:816-842) on the host mirror + device graph.
//   pass 1 (vacuum.rs:118-216)  dead heap TIDs removed; elements left with none form the skip set
//   pass 2 (vacuum.rs:413-640)  every live element that points at one of them, or whose layer-0 list is not full
//                               (vacuum.rs:228-281), gets all of its lists re-found: skip-set searches on the GPU in
//                               speculative batches (hnsw_repair_search_kernel), committed in element order.  A repair
//                               rewrites only the element's OWN lists, so a later search of the same batch is the one a
//                               sequential run would have made unless it expanded an element repaired earlier in the
//                               batch (the kernel's expansion log says so); then the round stops there and the rest is
//                               searched again.
//   pass 3 (vacuum.rs:655-790)  emptied elements are marked deleted, their lists and values cleared
// ------------------------------------------------------------------------------------------------
int pgvrx_index_vacuum(pgvrx_index *ix, const uint64_t *dead_tids, int64_t n_dead, int32_t host_threads, pgvrx_vacuum_stats *stats) try {
    if (!ix) return PGVRX_ERR_INVALID;
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_dead < 0 || (n_dead > 0 && !dead_tids)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null dead TID list");
    if (!ix->loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no graph loaded");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    CU_TRY(ix, cudaDeviceSynchronize());              // scans in flight read the lists this call rewrites
    ix->generation++;
    int avail = (int)std::max(1u, std::thread::hardware_concurrency());
    {
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) avail = std::min(avail, (int)CPU_COUNT(&set));
    }
    const int nthreads = host_threads > 0 ? host_threads : avail;
    if (!ix->build) {
        int rc = mirror_from_device(ix, nthreads);
        if (rc) return rc;
    }
    BuildState &g = *ix->build;
    const double t_start = now_ms();
    const int m = g.m, efc = g.efc, l0 = 2 * m;
    const int64_t n = g.n;

    // ---- pass 1 ----
    std::vector<uint64_t> dead(dead_tids, dead_tids + n_dead);
    std::sort(dead.begin(), dead.end());
    std::vector<char> deleted((size_t)n, 0);
    bool any_deleted = false;
    uint64_t removed = 0;
    int64_t hp = -1;
    int hp_level = -1;
    for (int64_t i = 0; i < n; i++) {
        uint8_t &cnt = g.tid_cnt[(size_t)i];
        uint64_t *t = &g.tids[(size_t)i * PGVRX_HEAPTIDS];
        if (cnt > 0) {
            int keep = 0;
            for (int k = 0; k < cnt; k++) {
                if (std::binary_search(dead.begin(), dead.end(), t[k])) removed++;
                else t[keep++] = t[k];
            }
            for (int k = keep; k < PGVRX_HEAPTIDS; k++) t[k] = ~0ull;
            cnt = (uint8_t)keep;
        }
        if (cnt == 0) { deleted[(size_t)i] = 1; any_deleted = true; }
        else if ((int)g.level[(size_t)i] > hp_level && i != g.entry) { hp = i; hp_level = g.level[(size_t)i]; }
    }
    if (stats) stats->n_heap_tids_removed = removed;
    if (!any_deleted) { if (stats) stats->total_ms = now_ms() - t_start; return PGVRX_OK; }

    // the kernels' view of the skip set and of elements deleted by earlier vacuums
    std::vector<uint8_t> flags((size_t)n);
    auto upload_flags = [&]() -> int {
        for (int64_t i = 0; i < n; i++)
            flags[(size_t)i] = (uint8_t)((g.del[(size_t)i] ? PGVRX_ELEM_DELETED : 0) | (g.tid_cnt[(size_t)i] == 0 ? PGVRX_ELEM_NO_HEAPTIDS : 0));
        CU_TRY(ix, ix->flags.ensure(std::max<size_t>(16, (size_t)ix->desc.capacity)));
        CU_TRY(ix, cudaMemcpy(ix->flags.p, flags.data(), (size_t)n, cudaMemcpyHostToDevice));
        ix->has_flags = true;
        return PGVRX_OK;
    };
    int rc = upload_flags();
    if (rc) return rc;

    // ---- pass 2 ----
    auto needs_updated = [&](int64_t i) -> bool {                 // vacuum.rs:228-281
        for (int lc = 0; lc <= g.level[(size_t)i]; lc++) {
            const int32_t *ids = g.ids(i, lc);
            const int c = g.cnt(i, lc);
            for (int t = 0; t < c; t++) if (deleted[(size_t)ids[t]]) return true;
        }
        return g.cnt0[(size_t)i] < l0;
    };
    if (!g.scratch) g.scratch = new RoundScratch();
    RoundBuffers &rb = g.scratch->rb;
    PinnedBuf &h_pack = g.scratch->h_pack;
    const int ef = efc + 1;                                       // insert.rs:1080-1084
    const int tail_cap = std::max(64, 1023 - ef);                 // room for skip members and boundary ties
    const int w_pitch = ef + tail_cap;
    const int rs_cap = std::max(256, 6 * ef + 64);
    const int rs_ent = 3 + (l0 + 31) / 32;
    uint64_t n_rounds = 0, n_searched = 0, n_conflicts = 0, n_repaired = 0;
    double kernel_ms = 0.0;
    std::vector<int64_t> sc_off0, sc_offu;
    std::vector<int32_t> sc_rows0, sc_rowsu;
    auto queue_row = [&](int64_t v, int lc) {
        if (lc == 0) {
            sc_off0.push_back(v * l0);
            sc_rows0.insert(sc_rows0.end(), &g.adj0[(size_t)v * l0], &g.adj0[(size_t)v * l0] + l0);
        } else {
            int64_t r = g.upper_base[(size_t)v] + lc - 1;
            sc_offu.push_back(r * m);
            sc_rowsu.insert(sc_rowsu.end(), &g.adj_up[(size_t)r * m], &g.adj_up[(size_t)r * m] + m);
        }
    };
    auto flush_scatter = [&]() -> int {
        for (int pass = 0; pass < 2; pass++) {
            std::vector<int64_t> &off = pass == 0 ? sc_off0 : sc_offu;
            std::vector<int32_t> &dat = pass == 0 ? sc_rows0 : sc_rowsu;
            if (off.empty()) continue;
            const int row_len = pass == 0 ? l0 : m;
            CU_TRY(ix, rb.sc_off.ensure(off.size() * 8));
            CU_TRY(ix, rb.sc_rows.ensure(dat.size() * 4));
            CU_TRY(ix, cudaMemcpyAsync(rb.sc_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, ix->stream));
            CU_TRY(ix, cudaMemcpyAsync(rb.sc_rows.p, dat.data(), dat.size() * 4, cudaMemcpyHostToDevice, ix->stream));
            CU_TRY(ix, launch_scatter_rows((int32_t *)(pass == 0 ? ix->adj0.p : ix->adj_up.p), (const int64_t *)rb.sc_off.p,
                                           (const int32_t *)rb.sc_rows.p, row_len, (int)off.size(), ix->stream));
            CU_TRY(ix, cudaStreamSynchronize(ix->stream));
            off.clear();
            dat.clear();
        }
        return PGVRX_OK;
    };

    // one GPU round: searches for `batch` from `entry`; commits a prefix in order; returns how many were committed
    std::vector<int32_t> h_elem;
    auto repair_round = [&](const std::vector<int64_t> &batch, int64_t entry, int *n_committed) -> int {
        const int B = (int)batch.size();
        *n_committed = 0;
        g.round++;
        n_rounds++;
        n_searched += (uint64_t)B;
        int max_layers = 1;
        for (int64_t e : batch) max_layers = std::max(max_layers, (int)g.level[(size_t)e] + 1);
        ScanConfig cfg;
        const uint32_t vis_entries = pgvrx_pow2_at_least(std::min<uint64_t>((uint64_t)ef * 2 * m * 4 * 16, std::max<uint64_t>(1024, (uint64_t)n * 2)));
        int rcc = pgvrx_make_config(ix, ef, tail_cap, vis_entries, B, ix->sm_count, true, &cfg);
        if (rcc) return rcc;
        int occ = 0;
        CU_TRY(ix, repair_kernel_occupancy(ix->desc.elem_type, ix->desc.metric, cfg.smem_block, cfg.wpb * 32, &occ));
        if (occ < 1) return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "repair search kernel does not fit on an SM");
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((B + cfg.wpb - 1) / cfg.wpb, (int64_t)ix->sm_count * occ));
        CU_TRY(ix, rb.vis.ensure((size_t)grid * cfg.wpb * cfg.vis_entries * 4));
        CU_TRY(ix, rb.elem_ids.ensure((size_t)B * 4));
        CU_TRY(ix, rb.work.ensure(64));
        size_t po = 0;
        auto carve = [&](size_t bytes) { const size_t o = po; po += (bytes + 255) & ~(size_t)255; return o; };
        const size_t o_wids = carve((size_t)B * max_layers * w_pitch * 4), o_wd = carve((size_t)B * max_layers * w_pitch * 4);
        const size_t o_wcnt = carve((size_t)B * max_layers * 4), o_rs = carve((size_t)B * rs_cap * rs_ent * 4);
        const size_t o_rscnt = carve((size_t)B * 4), o_status = carve((size_t)B * 4), o_ctr = carve((size_t)B * 16);
        CU_TRY(ix, rb.out_pack.ensure(po));
        CU_TRY(ix, h_pack.ensure(po));
        uint8_t *dpk = (uint8_t *)rb.out_pack.p;
        const uint8_t *hpk = (const uint8_t *)h_pack.p;
        h_elem.resize((size_t)B);
        for (int i = 0; i < B; i++) h_elem[(size_t)i] = (int32_t)batch[(size_t)i];
        CU_TRY(ix, cudaMemcpyAsync(rb.elem_ids.p, h_elem.data(), (size_t)B * 4, cudaMemcpyHostToDevice, ix->stream));
        CU_TRY(ix, cudaMemsetAsync(rb.work.p, 0, 64, ix->stream));
        ix->n = n;
        ix->entry = entry;
        ix->entry_level = entry >= 0 ? g.level[(size_t)entry] : -1;
        GraphView gv = pgvrx_view_of(ix);
        RepairParams rp;
        rp.elem_ids = (const int32_t *)rb.elem_ids.p;
        rp.n_pending = B;
        rp.ef = ef;
        rp.tail_cap = tail_cap;
        rp.max_layers = max_layers;
        rp.w_pitch = w_pitch;
        rp.w_ids = (int32_t *)(dpk + o_wids);
        rp.w_dists = (float *)(dpk + o_wd);
        rp.w_counts = (int32_t *)(dpk + o_wcnt);
        rp.rs_log = (int32_t *)(dpk + o_rs);
        rp.rs_cap = rs_cap;
        rp.rs_ent = rs_ent;
        rp.rs_counts = (int32_t *)(dpk + o_rscnt);
        rp.out_status = (uint32_t *)(dpk + o_status);
        rp.out_ctr = (uint32_t *)(dpk + o_ctr);
        rp.vis_base = (uint32_t *)rb.vis.p;
        rp.vis_entries = cfg.vis_entries;
        rp.up_entries = cfg.vis_entries;
        rp.work_counter = (int *)rb.work.p;
        rp.stage_stride = cfg.L.stage_stride;
        rp.R = cfg.L.R;
        rp.stage_bytes = cfg.L.stage_bytes;
        rp.chunk_bytes = cfg.L.chunk_bytes;
        rp.chunk_stride = cfg.L.chunk_stride;
        CU_TRY(ix, cudaEventRecord(ix->ev0, ix->stream));
        CU_TRY(ix, launch_repair_search(ix->desc.elem_type, ix->desc.metric, gv, rp, grid, cfg.smem_block, cfg.wpb * 32, ix->stream));
        CU_TRY(ix, cudaEventRecord(ix->ev1, ix->stream));
        CU_TRY(ix, cudaMemcpyAsync(h_pack.p, rb.out_pack.p, po, cudaMemcpyDeviceToHost, ix->stream));
        CU_TRY(ix, cudaStreamSynchronize(ix->stream));
        {
            float kms = 0.0f;
            CU_TRY(ix, cudaEventElapsedTime(&kms, ix->ev0, ix->ev1));
            kernel_ms += kms;
        }
        const int32_t *h_wids = (const int32_t *)(hpk + o_wids);
        const float *h_wd = (const float *)(hpk + o_wd);
        const int32_t *h_wcnt = (const int32_t *)(hpk + o_wcnt);
        const int32_t *h_rs = (const int32_t *)(hpk + o_rs);
        const int32_t *h_rscnt = (const int32_t *)(hpk + o_rscnt);
        const uint32_t *h_status = (const uint32_t *)(hpk + o_status);
        const int entry_level = entry >= 0 ? g.level[(size_t)entry] : -1;
        for (int j = 0; j < B; j++) {
            const int64_t e = batch[(size_t)j];
            if (h_status[(size_t)j] != 0)
                return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "vacuum repair search ran out of scratch (skip set too dense around an element)");
            if (j > 0) {                                          // did it expand an element repaired earlier in this round?
                bool valid = h_rscnt[(size_t)j] <= rs_cap;
                const int32_t *rs = &h_rs[(size_t)j * rs_cap * rs_ent];
                for (int t = 0; valid && t < h_rscnt[(size_t)j]; t++)
                    if (g.stamp[(size_t)rs[(size_t)rs_ent * t]] == g.round) valid = false;
                if (!valid) { n_conflicts++; break; }
            }
            // vacuum.rs:366-377: every layer's list := the nearest lm of W that are not in the skip set (insert.rs:1101-1116)
            const int lv = g.level[(size_t)e];
            for (int lc = 0; lc <= lv; lc++) {
                int32_t *ids = g.ids(e, lc);
                float *ds = g.dists(e, lc);
                const int lm = g.lm(lc);
                int c = 0;
                if (lc <= entry_level && lc < max_layers) {
                    const int cntw = h_wcnt[(size_t)j * max_layers + lc];
                    const size_t o = ((size_t)j * max_layers + lc) * w_pitch;
                    for (int t = 0; t < cntw && c < lm; t++) {
                        const int32_t x = h_wids[o + t];
                        if (x == (int32_t)e || deleted[(size_t)x]) continue;
                        ids[c] = x;
                        ds[c] = h_wd[o + t];
                        c++;
                    }
                }
                for (int t = c; t < lm; t++) ids[t] = -1;
                g.cnt(e, lc) = (uint8_t)c;
                queue_row(e, lc);
            }
            g.stamp[(size_t)e] = g.round;
            n_repaired++;
            (*n_committed)++;
            if (lv > entry_level) {                               // vacuum.rs:614-623: it becomes the entry point
                g.entry = e;
                g.entry_level = lv;
                break;
            }
        }
        return flush_scatter();
    };

    // repair_graph_entry_point (vacuum.rs:413-523)
    int done = 0;
    if (hp >= 0 && !g.del[(size_t)hp] && needs_updated(hp) && g.entry >= 0) {
        // (the highest point never outranks the entry point here: the entry-level update below applies to the main loop only)
        const int64_t save_entry = g.entry;
        const int save_level = g.entry_level;
        rc = repair_round(std::vector<int64_t>{hp}, g.entry, &done);
        if (rc) return rc;
        g.entry = save_entry;
        g.entry_level = save_level;
    }
    if (g.entry >= 0) {
        if (deleted[(size_t)g.entry]) {                           // vacuum.rs:464-484
            g.entry = hp;
            g.entry_level = hp >= 0 ? g.level[(size_t)hp] : -1;
        } else if (!g.del[(size_t)g.entry] && needs_updated(g.entry) && hp >= 0) {   // vacuum.rs:486-514 (without a highest point
            const int64_t save_entry = g.entry;                                      //  the element is its own entry: no-op)
            const int save_level = g.entry_level;
            rc = repair_round(std::vector<int64_t>{g.entry}, hp, &done);
            if (rc) return rc;
            g.entry = save_entry;
            g.entry_level = save_level;
        }
    }
    // every other live element, in element (= page) order (vacuum.rs:549-640)
    int window = 8;
    int64_t pos = 0;
    std::vector<int64_t> batch;
    while (pos < n) {
        batch.clear();
        int64_t scan = pos;
        while (scan < n && (int)batch.size() < window) {
            const int64_t i = scan++;
            if (g.tid_cnt[(size_t)i] == 0 || g.del[(size_t)i] || i == g.entry) continue;      // vacuum.rs:566-570, :300-302
            if (!needs_updated(i)) continue;
            batch.push_back(i);
        }
        if (batch.empty()) break;
        if (g.entry < 0) return pgvrx_fail(ix, PGVRX_ERR_STATE, "live elements but no entry point");
        rc = repair_round(batch, g.entry, &done);
        if (rc) return rc;
        // elements before the first uncommitted one are finished; the rest is looked at again
        pos = done < (int)batch.size() ? batch[(size_t)done] : scan;
        window = std::min(512, std::max(8, 2 * done + 2));
    }

    // ---- pass 3 (vacuum.rs:655-790) ----
    std::vector<uint8_t> zero_row(ix->row_stride, 0);
    uint64_t n_marked = 0;
    for (int64_t i = 0; i < n; i++) {
        if (g.del[(size_t)i] || g.tid_cnt[(size_t)i] != 0) continue;
        for (int lc = 0; lc <= g.level[(size_t)i]; lc++) {
            int32_t *ids = g.ids(i, lc);
            for (int t = 0; t < g.lm(lc); t++) ids[t] = -1;
            g.cnt(i, lc) = 0;
            queue_row(i, lc);
        }
        g.del[(size_t)i] = 1;
        memset(g.values.data() + (size_t)i * g.row_bytes, 0, g.row_bytes);
        CU_TRY(ix, cudaMemcpyAsync((uint8_t *)ix->vectors.p + (size_t)i * ix->row_stride, zero_row.data(), ix->row_stride,
                                   cudaMemcpyHostToDevice, ix->stream));
        n_marked++;
    }
    rc = flush_scatter();
    if (rc) return rc;
    CU_TRY(ix, cudaStreamSynchronize(ix->stream));
    rc = upload_flags();
    if (rc) return rc;
    ix->n = n;
    ix->entry = g.entry;
    ix->entry_level = g.entry_level;
    ix->tc_ready = false;                                         // rows of deleted elements were zeroed
    if (stats) {
        stats->n_elements_deleted = n_marked;
        stats->n_elements_repaired = n_repaired;
        stats->n_rounds = n_rounds;
        stats->n_searched = n_searched;
        stats->n_conflicts = n_conflicts;
        stats->kernel_ms = kernel_ms;
        stats->total_ms = now_ms() - t_start;
    }
    return PGVRX_OK;
} catch (const std::bad_alloc &) {
    return host_oom(ix);
}

int pgvrx_index_element_flags(pgvrx_index *ix, uint8_t *flags, int64_t n) {
    if (!ix || !flags) return PGVRX_ERR_INVALID;
    if (!ix->loaded || n != ix->n) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "flags must cover every element of the index");
    if (!ix->has_flags) { memset(flags, 0, (size_t)n); return PGVRX_OK; }
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    CU_TRY(ix, cudaMemcpy(flags, ix->flags.p, (size_t)n, cudaMemcpyDeviceToHost));
    return PGVRX_OK;
}

int pgvrx_index_set_heap_tids(pgvrx_index *ix, const uint64_t *heap_tids, const uint8_t *counts, int64_t n) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (!ix->loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no graph loaded");
    if (!heap_tids || !counts || n != ix->n) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "heap TID arrays must cover every element of the index");
    for (int64_t i = 0; i < n; i++)
        if (counts[i] > PGVRX_HEAPTIDS) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "an element holds at most 10 heap TIDs");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    if (!ix->build) {
        int avail = (int)std::max(1u, std::thread::hardware_concurrency());
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) avail = std::min(avail, (int)CPU_COUNT(&set));
        int rc = mirror_from_device(ix, avail);
        if (rc) return rc;
    }
    BuildState &g = *ix->build;
    memcpy(g.tids.data(), heap_tids, (size_t)n * PGVRX_HEAPTIDS * 8);
    memcpy(g.tid_cnt.data(), counts, (size_t)n);
    return PGVRX_OK;
}

int pgvrx_index_graph_info(const pgvrx_index *ix, int64_t *n, int64_t *n_upper_rows, int64_t *entry, int32_t *entry_level) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (n) *n = ix->n;
    if (n_upper_rows) *n_upper_rows = ix->n_upper;
    if (entry) *entry = ix->entry;
    if (entry_level) *entry_level = ix->entry_level;
    return PGVRX_OK;
}

int pgvrx_index_export_graph(pgvrx_index *ix, void *rows, uint8_t *level, int32_t *adj0, int32_t *upper_base, int32_t *adj_up,
                             uint64_t *heap_tids, uint8_t *heap_tid_counts) {
    if (!ix) return PGVRX_ERR_INVALID;
    if (!ix->build) {
        // a handle that was loaded (or fed through pgvrx_commit_insert) has no host mirror yet: read the arrays back
        if (!ix->loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no graph to export");
        CU_TRY(ix, cudaSetDevice(ix->desc.device));
        int avail = (int)std::max(1u, std::thread::hardware_concurrency());
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) avail = std::min(avail, (int)CPU_COUNT(&set));
        try {
            int rc = mirror_from_device(ix, avail);
            if (rc) return rc;
        } catch (const std::bad_alloc &) {
            return host_oom(ix);
        }
    }
    BuildState &g = *ix->build;
    const int m = g.m;
    if (rows) memcpy(rows, g.values.data(), (size_t)g.n * g.row_bytes);
    if (level) memcpy(level, g.level.data(), (size_t)g.n);
    if (adj0) memcpy(adj0, g.adj0.data(), (size_t)g.n * 2 * m * 4);
    if (upper_base) memcpy(upper_base, g.upper_base.data(), (size_t)g.n * 4);
    if (adj_up) memcpy(adj_up, g.adj_up.data(), (size_t)g.n_upper * m * 4);
    if (heap_tids) memcpy(heap_tids, g.tids.data(), (size_t)g.n * PGVRX_HEAPTIDS * 8);
    if (heap_tid_counts) memcpy(heap_tid_counts, g.tid_cnt.data(), (size_t)g.n);
    return PGVRX_OK;
}

}  // extern "C"
