// hnsw_device.cuh -- warp-level building blocks of the B200 HNSW search kernels.
//
// One warp owns one query.  Per expansion the warp
//   1. reads the candidate's neighbour row (one coalesced 128-B load for m = 16),
//   2. tests/marks every neighbour in an exact per-query visited hash table (global, L2 resident),
//   3. gathers the unvisited neighbours' vectors into shared memory with one TMA bulk copy
//      (cp.async.bulk, mbarrier completion) per row,
//   4. evaluates one distance per LANE, each lane walking its row strictly left to right with
//      separate multiply and add -- bit-identical to the reference's scalar loops
//      (src/types/vector.rs:518-567, halfvec.rs:687-733, bitvec.rs:97-132),
//   5. applies the reference's accept / evict rule in neighbour order
//      (src/graph/mod.rs:205-243, src/index/scan.rs:362-430) against a sorted result list.
//
// The candidate heap C and the result heap W of the reference are one sorted array S of
// (distance, id) keys: W = S[0, ef), an "expanded" flag per entry replaces C, and entries evicted
// from W that still tie with W's furthest distance are retained behind position ef because the
// reference would still expand them (stop rule is strict `>`, graph/mod.rs:191, scan.rs:340).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hnsw_kernels.h"

namespace pgvrx {

constexpr uint32_t VIS_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t FULL = 0xFFFFFFFFu;

enum : int { ELEM_F32 = 0, ELEM_F16 = 1, ELEM_BIT = 2 };
enum : int { METRIC_L2SQ = 0, METRIC_NEG_IP = 1, METRIC_L1 = 2, METRIC_HAMMING = 3, METRIC_JACCARD = 4 };

// per-query status bits (non-zero => the host re-runs the query with larger scratch)
enum : uint32_t { ST_VIS_OVERFLOW = 1u, ST_TIE_OVERFLOW = 2u, ST_POOL_OVERFLOW = 4u };


// ------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA bulk copy (global -> shared), L2 eviction policy
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// TMA 1-D bulk copy of one vector row: SASS UBLKCP.  dst/src 16-B aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_row_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// Ampere-style async copy, 16 B per lane (SASS LDGSTS.E.BYPASS.128): kept as a build-time alternative
// (-DPGVRX_GATHER_CPASYNC) for the row gather.  Measured on B200 it is ~4x SLOWER than per-row TMA bulk
// copies for this access pattern (12 warps x 25 rows in flight per SM queue up in L1TEX, which TMA
// bypasses): 7990 vs 2060 cycles per expansion, profiles/r1_notes.md.
__device__ __forceinline__ void cp_async16(void *dst, const void *src, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// keys: (order-preserving f32 bits) << 32 | id << 1 | expanded
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f32_ord(float d) {
    d = __fadd_rn(d, 0.0f);  // -0.0 -> +0.0 : the reference compares with partial_cmp, where they are equal
    uint32_t b = __float_as_uint(d);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_f32(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
    return ((uint64_t)f32_ord(d) << 32) | ((uint64_t)id << 1);
}
__device__ __forceinline__ uint32_t key_ord(uint64_t k) { return (uint32_t)(k >> 32); }
__device__ __forceinline__ uint32_t key_id(uint64_t k) { return ((uint32_t)k) >> 1; }

// ------------------------------------------------------------------------------------------
// exact visited set: open addressing in a per-warp global table (HashSet semantics,
// graph/mod.rs:171,206-209; scan.rs:317,366-369)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vis_hash(uint32_t id) { return id * 0x9E3779B1u; }

// Layout: linear probing over the whole table, scanned in 16-byte buckets of 4 slots.  An id lives in
// the first slot that was free at/after its home slot 4*bucket(id) when it was inserted; slots never
// become free again within a search, so "the home bucket still has a free slot and does not hold the
// id" proves absence with ONE 128-bit load.
__device__ __forceinline__ uint32_t vis_home(uint32_t id, uint32_t mask) { return ((vis_hash(id) >> 7) & mask) & ~3u; }

// synchronous test-and-set starting at slot `h`: returns true when id was not in the set (and is now)
__device__ __forceinline__ bool vis_insert_from(uint32_t *tab, uint32_t mask, uint32_t h, uint32_t id) {
    for (uint32_t probes = 0; probes <= mask; probes++) {
        uint32_t old = atomicCAS(tab + h, VIS_EMPTY, id);
        if (old == VIS_EMPTY) return true;
        if (old == id) return false;
        h = (h + 1) & mask;
    }
    return false;  // unreachable: the caller keeps the load factor <= 3/4
}
__device__ __forceinline__ bool vis_insert(uint32_t *tab, uint32_t mask, uint32_t id) {
    return vis_insert_from(tab, mask, vis_home(id, mask), id);
}
__device__ __forceinline__ void vis_clear(uint32_t *tab, uint32_t entries, int lane) {
    uint4 e = make_uint4(VIS_EMPTY, VIS_EMPTY, VIS_EMPTY, VIS_EMPTY);
    uint4 *t4 = reinterpret_cast<uint4 *>(tab);
    for (uint32_t i = lane; i < entries / 4; i += 32) t4[i] = e;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// distances: one LANE per row, strict left-to-right f32 accumulation, no FMA contraction
// ------------------------------------------------------------------------------------------
// Blackwell packed fp32: sub.rn.f32x2 / mul.rn.f32x2 (SASS FADD2 / FMUL2) do two IEEE-exact
// element operations per instruction; only the accumulation has to stay a scalar chain.
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    uint64_t ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    uint64_t ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

// two consecutive elements: element-wise part packed, accumulation strictly in element order
template <int METRIC>
__device__ __forceinline__ float acc_step2(float acc, float2 q, float2 x) {
    if (METRIC == METRIC_L2SQ) {
        float2 d = sub2(q, x);                          // vector.rs:522
        float2 sq = mul2(d, d);
        acc = __fadd_rn(acc, sq.x);                     // vector.rs:523
        return __fadd_rn(acc, sq.y);
    } else if (METRIC == METRIC_NEG_IP) {
        float2 pr = mul2(q, x);                         // vector.rs:534
        acc = __fadd_rn(acc, pr.x);
        return __fadd_rn(acc, pr.y);
    } else {
        float2 d = sub2(q, x);                          // vector.rs:564
        acc = __fadd_rn(acc, fabsf(d.x));
        return __fadd_rn(acc, fabsf(d.y));
    }
}

// Resumable form of one (query, row) distance: the row may arrive in several chunks (wide rows are staged 512 bytes at a
// time so that all 32 lanes of a warp have a row to work on); chunks are added strictly in element order, so the result
// is the same chain of operations as a single pass -- bit-identical to the reference's scalar loops.
// q, row: shared memory, 16-B aligned, nvec 16-byte vectors (zero padded: padding adds +0.0 exactly)
template <int ELEM, int METRIC>
struct DistAcc {
    float acc;
    uint32_t x, ab, aa, bb;
    __device__ __forceinline__ void init() { acc = 0.0f; x = ab = aa = bb = 0; }
    __device__ __forceinline__ void add(const uint8_t *q, const uint8_t *row, int nvec) {
        const uint4 *q4 = reinterpret_cast<const uint4 *>(q);
        const uint4 *r4 = reinterpret_cast<const uint4 *>(row);
        if (ELEM == ELEM_F32) {
            float a_ = acc;
#pragma unroll 4
            for (int j = 0; j < nvec; j++) {
                uint4 a = q4[j], b = r4[j];
                a_ = acc_step2<METRIC>(a_, make_float2(__uint_as_float(a.x), __uint_as_float(a.y)),
                                       make_float2(__uint_as_float(b.x), __uint_as_float(b.y)));
                a_ = acc_step2<METRIC>(a_, make_float2(__uint_as_float(a.z), __uint_as_float(a.w)),
                                       make_float2(__uint_as_float(b.z), __uint_as_float(b.w)));
            }
            acc = a_;
        } else if (ELEM == ELEM_F16) {
            float a_ = acc;
#pragma unroll 2
            for (int j = 0; j < nvec; j++) {
                uint4 a = q4[j], b = r4[j];
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    // half -> f32 is exact for every value (halfvec.rs:54-87 restates IEEE decoding)
                    float2 fa = __half22float2(*reinterpret_cast<const __half2 *>(&aw[t]));
                    float2 fb = __half22float2(*reinterpret_cast<const __half2 *>(&bw[t]));
                    a_ = acc_step2<METRIC>(a_, fa, fb);
                }
            }
            acc = a_;
        } else {
            // bitvec.rs:97-132: integer popcounts (order-free, exact)
#pragma unroll 4
            for (int j = 0; j < nvec; j++) {
                uint4 a = q4[j], b = r4[j];
                if (METRIC == METRIC_HAMMING) {
                    x += __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
                } else {
                    ab += __popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w);
                    aa += __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w);
                    bb += __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
                }
            }
        }
    }
    __device__ __forceinline__ float finish() const {
        if (ELEM == ELEM_BIT) {
            if (METRIC == METRIC_HAMMING) return (float)x;
            if (ab == 0) return 1.0f;                                     // bitvec.rs:127-128
            double j = 1.0 - ((double)ab / (double)(aa + bb - ab));       // bitvec.rs:130
            return (float)j;                                              // build.rs:367 `as f32`
        }
        return METRIC == METRIC_NEG_IP ? -acc : acc;                      // vector.rs:631 (f64 negate, exact)
    }
};

template <int ELEM, int METRIC>
__device__ __forceinline__ float row_distance(const uint8_t *q, const uint8_t *row, int nvec) {
    DistAcc<ELEM, METRIC> a;
    a.init();
    a.add(q, row, nvec);
    return a.finish();
}

// ------------------------------------------------------------------------------------------
// per-warp search state
// ------------------------------------------------------------------------------------------
struct WarpCtx {
    uint8_t *stage;        // R rows, stage_stride apart
    uint8_t *qsm;          // query row
    int32_t *nbr;          // 32 compacted neighbour ids
    uint64_t *mbar;
    uint32_t *vis;         // global visited table of this warp slot
    uint64_t policy;
    uint32_t stage_stride;
    int R, nvec;
    uint32_t chunk_bytes;  // 0: whole rows are staged (R at a time); else rows are staged in chunks of this many bytes,
    uint32_t chunk_stride; //    32 rows at a time, chunk_stride bytes apart (wide rows: every lane keeps a row to work on)
    uint8_t *stage2;       // chunked staging with room for two chunk buffers: the second one (else NULL); chunk i + 1 is then
    uint32_t phase2;       //    in flight (second mbarrier, c.mbar + 1) while chunk i is accumulated
    int len, ef, cap;      // list length (W plus retained boundary ties), ef, capacity
    uint32_t phase;
    uint32_t f_ord;        // ordered bits of W's furthest distance (valid when len >= ef)
    uint32_t n_dist, n_exp0, n_expu, n_vis, status;
    int32_t *rs_log;       // build: 3 ints per expansion -- node id, ordered bits of f_dist (or -1), layer (or NULL)
    int rs_cap, rs_n;      // capacity / count (rs_n may exceed rs_cap: the host then treats the set as unknown)
    int rs_ent;            // ints per log entry: 3 + one word per 32 list slots
    bool spec;             // fetch every neighbour's row while the visited probes are in flight (see search_layer)
    bool hybrid;           // whole-row staging in several rounds: round 0 speculative (issued before the probes resolve), the rest probe-first
    int32_t self_id;       // vacuum repair search only: the element being repaired (member of the skip set)
    int excess;            // vacuum repair search only: counted entry points seeded beyond ef (w_len - ef, scan.rs:326-336)
    int lane;
#ifdef PGVRX_PHASE_TIMING
    uint32_t t[8];         // cycles: 0 pop+adjacency, 1 visited, 2 tma issue+wait, 3 distance, 4 accept/insert
    long long t_last;
#endif
};

#ifdef PGVRX_PHASE_TIMING
#define PH_START(c) (c).t_last = clock64()
#define PH_TICK(c, slot) do { long long _n = clock64(); (c).t[slot] += (uint32_t)(_n - (c).t_last); (c).t_last = _n; } while (0)
#else
#define PH_START(c) do {} while (0)
#define PH_TICK(c, slot) do {} while (0)
#endif

constexpr uint64_t KEY_NONE = ~0ull;

// ---- S in registers: entry i lives in lane i / J, register i % J (J * 32 >= cap + 1) -------------
// No shared-memory round trips: position by REDUX, shift by one shuffle, flag test by ballot.
template <int J>
struct RegList {
    uint64_t k[J];

    __device__ __forceinline__ uint64_t get(const WarpCtx &c, int i) const {
        uint64_t v = 0;
#pragma unroll
        for (int j = 0; j < J; j++) if ((i % J) == j) v = k[j];
        return __shfl_sync(FULL, v, i / J);
    }
    __device__ __forceinline__ void reset1(WarpCtx &c, uint64_t key) {
#pragma unroll
        for (int j = 0; j < J; j++) k[j] = KEY_NONE;
        if (c.lane == 0) k[0] = key;
        c.len = 1;
        if (c.ef == 1) c.f_ord = key_ord(key);
    }
    __device__ __forceinline__ void insert(WarpCtx &c, uint64_t key) {
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < J; j++) cnt += k[j] < key;           // KEY_NONE is never < key
        const int pos = __reduce_add_sync(FULL, cnt);
        const uint64_t carry = __shfl_up_sync(FULL, k[J - 1], 1);
#pragma unroll
        for (int j = J - 1; j >= 0; j--) {
            const int gi = c.lane * J + j;
            const uint64_t src = j > 0 ? k[j > 0 ? j - 1 : 0] : carry;
            k[j] = gi > pos ? src : (gi == pos ? key : k[j]);
        }
        c.len += 1;
        if (c.len >= c.ef) {
            const uint32_t old_f = c.f_ord;
            c.f_ord = key_ord(get(c, c.ef - 1));
            if (c.len > c.ef) {
                // Entries behind W all tie with the OLD furthest distance (invariant).  The element just
                // pushed out of W has that distance too, so: if W's furthest distance is unchanged they all
                // stay (the reference would still expand them, stop rule is strict `>`); otherwise none can.
                if (c.f_ord == old_f) {
                    if (c.len > c.cap) { c.len = c.cap; c.status |= ST_TIE_OVERFLOW; }
                } else {
                    c.len = c.ef;
                }
#pragma unroll
                for (int j = 0; j < J; j++) if (c.lane * J + j >= c.len) k[j] = KEY_NONE;
            }
        }
    }
    // nearest unexpanded entry: mark it expanded and return its key; false when none is left
    __device__ __forceinline__ bool pop_nearest(WarpCtx &c, uint64_t &key) {
        int jf = -1;
#pragma unroll
        for (int j = J - 1; j >= 0; j--) if (k[j] != KEY_NONE && (k[j] & 1ull) == 0) jf = j;
        const uint32_t b = __ballot_sync(FULL, jf >= 0);
        if (!b) return false;
        const int src = __ffs(b) - 1;
        uint64_t mine = 0;
#pragma unroll
        for (int j = 0; j < J; j++) if (j == jf) mine = k[j];
        key = __shfl_sync(FULL, mine, src);
        if (c.lane == src) {
#pragma unroll
            for (int j = 0; j < J; j++) if (j == jf) k[j] |= 1ull;
        }
        return true;
    }
    // Build phase 2 (graph/mod.rs:425 `ep = w`): W of this layer becomes the entry-point list of the next
    // one.  Drops retained ties (they are not in W), clears the expanded flags, sets f_ord, and inserts
    // every id into the (freshly cleared) visited table.
    __device__ __forceinline__ void reopen(WarpCtx &c, uint32_t *vis, uint32_t vis_mask) {
        c.len = min(c.len, c.ef);
#pragma unroll
        for (int j = 0; j < J; j++) {
            const int gi = c.lane * J + j;
            if (gi >= c.len) k[j] = KEY_NONE;
            else { k[j] &= ~1ull; vis_insert(vis, vis_mask, key_id(k[j])); }
        }
        __syncwarp();
        if (c.len >= c.ef) c.f_ord = key_ord(get(c, c.ef - 1));
    }
    __device__ __forceinline__ void store(const WarpCtx &c, int32_t *ids, float *dists, int out_n) const {
#pragma unroll
        for (int j = 0; j < J; j++) {
            const int gi = c.lane * J + j;
            if (gi < out_n) { ids[gi] = (int32_t)key_id(k[j]); dists[gi] = ord_f32(key_ord(k[j])); }
        }
    }
};

// ---- S in shared memory (any ef up to 1000) -------------------------------------------------------
struct SmemList {
    uint64_t *S;   // cap + 1 keys
    int hint;      // every entry below hint is expanded

    __device__ __forceinline__ uint64_t get(const WarpCtx &c, int i) const { return S[i]; }
    __device__ __forceinline__ void reset1(WarpCtx &c, uint64_t key) {
        __syncwarp();
        if (c.lane == 0) S[0] = key;
        __syncwarp();
        c.len = 1;
        hint = 0;
        if (c.ef == 1) c.f_ord = key_ord(key);
    }
    __device__ __forceinline__ void insert(WarpCtx &c, uint64_t key) {
        const int lane = c.lane;
        int pos = 0;
        for (int base = 0; base < c.len; base += 32) {
            int i = base + lane;
            bool lt = i < c.len && S[i] < key;
            pos += __popc(__ballot_sync(FULL, lt));
        }
        // shift [pos, len) right by one, highest chunk first
        for (int base = ((c.len - 1) >> 5) << 5; base >= 0 && base + 31 >= pos && c.len > 0; base -= 32) {
            int i = base + lane;
            bool mv = i >= pos && i < c.len;
            uint64_t v = mv ? S[i] : 0;
            __syncwarp();
            if (mv) S[i + 1] = v;
            __syncwarp();
        }
        if (lane == 0) S[pos] = key;
        __syncwarp();
        c.len += 1;
        if (pos < hint) hint = pos;
        if (c.len >= c.ef) {
            c.f_ord = key_ord(S[c.ef - 1]);
            if (c.len > c.ef) {
                int keep = 0;
                for (int base = c.ef; base < c.len; base += 32) {
                    int i = base + lane;
                    bool eq = i < c.len && key_ord(S[i]) == c.f_ord;
                    keep += __popc(__ballot_sync(FULL, eq));
                }
                c.len = c.ef + keep;
                if (c.len > c.cap) { c.len = c.cap; c.status |= ST_TIE_OVERFLOW; }
            }
        }
    }
    __device__ __forceinline__ bool pop_nearest(WarpCtx &c, uint64_t &key) {
        for (int base = (hint >> 5) << 5; base < c.len; base += 32) {
            int i = base + c.lane;
            bool un = i >= hint && i < c.len && (S[i] & 1ull) == 0;
            uint32_t b = __ballot_sync(FULL, un);
            if (b) {
                int idx = base + __ffs(b) - 1;
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
    __device__ __forceinline__ void reopen(WarpCtx &c, uint32_t *vis, uint32_t vis_mask) {
        c.len = min(c.len, c.ef);
        __syncwarp();
        for (int i = c.lane; i < c.len; i += 32) {
            S[i] &= ~1ull;
            vis_insert(vis, vis_mask, key_id(S[i]));
        }
        __syncwarp();
        hint = 0;
        if (c.len >= c.ef) c.f_ord = key_ord(S[c.ef - 1]);
    }
    __device__ __forceinline__ void store(const WarpCtx &c, int32_t *ids, float *dists, int out_n) const {
        for (int i = c.lane; i < out_n; i += 32) { ids[i] = (int32_t)key_id(S[i]); dists[i] = ord_f32(key_ord(S[i])); }
    }
};

// Gather nrow vector rows (ids in c.nbr[s .. s+nrow)) into the staging area and wait for them.
// Row r lands at c.stage + r * stage_stride.
__device__ __forceinline__ void gather_rows(WarpCtx &c, const GraphView &g, int s, int nrow) {
    const int lane = c.lane;
#ifndef PGVRX_GATHER_CPASYNC
    // one TMA bulk copy (UBLKCP) per row, issued by the lane that owns the row, one mbarrier phase
    int32_t my = lane < nrow ? c.nbr[s + lane] : -1;
    if (lane == 0) mbar_arrive_expect_tx(c.mbar, (uint32_t)nrow * g.row_stride);
    __syncwarp();
    if (lane < nrow)
        tma_row_g2s(c.stage + (size_t)lane * c.stage_stride, g.vectors + (size_t)my * g.row_stride, g.row_stride,
                    c.mbar, c.policy);
    mbar_wait(c.mbar, c.phase);
    c.phase ^= 1u;
#else
    const int nvec = c.nvec;
    if (nvec >= 32) {
        // wide rows: the whole warp copies one row per step, 512 contiguous bytes per instruction
        for (int r = 0; r < nrow; r++) {
            const uint8_t *src = g.vectors + (size_t)c.nbr[s + r] * g.row_stride;
            uint8_t *dst = c.stage + (size_t)r * c.stage_stride;
            for (int v = lane; v < nvec; v += 32) cp_async16(dst + v * 16, src + v * 16, c.policy);
        }
    } else {
        // narrow rows (e.g. 1024-bit vectors = 8 x 16 B): several rows per instruction
        const int rpi = 32 / nvec;                    // rows per instruction
        const int lr = lane / nvec, lc = lane - lr * nvec;
        for (int r0 = 0; r0 < nrow; r0 += rpi) {
            const int r = r0 + lr;
            if (lr < rpi && r < nrow)
                cp_async16(c.stage + (size_t)r * c.stage_stride + lc * 16,
                           g.vectors + (size_t)c.nbr[s + r] * g.row_stride + lc * 16, c.policy);
        }
    }
    cp_async_wait_all();
    __syncwarp();
#endif
}

// Same, but only the lanes with `mine` set (lane r <-> staging slot r - s) fetch their row.
__device__ __forceinline__ void gather_rows_masked(WarpCtx &c, const GraphView &g, int s, int nrow, bool mine) {
    const int lane = c.lane;
    const bool issue = mine && lane >= s && lane < s + nrow;
    const uint32_t m = __ballot_sync(FULL, issue);
    if (!m) return;
    if (lane == 0) mbar_arrive_expect_tx(c.mbar, (uint32_t)__popc(m) * g.row_stride);
    __syncwarp();
    if (issue)
        tma_row_g2s(c.stage + (size_t)(lane - s) * c.stage_stride, g.vectors + (size_t)c.nbr[lane] * g.row_stride, g.row_stride,
                    c.mbar, c.policy);
    mbar_wait(c.mbar, c.phase);
    c.phase ^= 1u;
}

// Bytes [off, off + cb) of the rows of the lanes with `mine` set (lane r <-> staging slot r, chunk_stride apart).
__device__ __forceinline__ void gather_chunk(WarpCtx &c, const GraphView &g, bool mine, uint32_t off, uint32_t cb) {
    const uint32_t m = __ballot_sync(FULL, mine);
    if (!m) return;
    if (c.lane == 0) mbar_arrive_expect_tx(c.mbar, (uint32_t)__popc(m) * cb);
    __syncwarp();
    if (mine)
        tma_row_g2s(c.stage + (size_t)c.lane * c.chunk_stride, g.vectors + (size_t)c.nbr[c.lane] * g.row_stride + off, cb, c.mbar, c.policy);
    mbar_wait(c.mbar, c.phase);
    c.phase ^= 1u;
}

// One distance per lane for the neighbours c.nbr[0 .. cnt) (lane r <-> neighbour r); lanes without `want` skip theirs.
// Whole-row staging: R rows per round, the lanes of the round evaluate.  Chunked staging (wide rows): all rows at once,
// 512 bytes at a time, every lane accumulating its own row's chain chunk after chunk.
template <int ELEM, int METRIC>
__device__ __forceinline__ float eval_neighbours(WarpCtx &c, const GraphView &g, int cnt, bool want, bool masked) {
    const int lane = c.lane;
    float d = 0.0f;
    if (c.chunk_bytes == 0) {
        for (int s = 0; s < cnt; s += c.R) {
            const int nrow = min(c.R, cnt - s);
            if (masked) gather_rows_masked(c, g, s, nrow, want);
            else gather_rows(c, g, s, nrow);
            PH_TICK(c, 2);
            if (want && lane >= s && lane < s + nrow)
                d = row_distance<ELEM, METRIC>(c.qsm, c.stage + (size_t)(lane - s) * c.stage_stride, c.nvec);
            __syncwarp();   // all lanes done with the staging rows before they are overwritten
            PH_TICK(c, 3);
        }
    } else {
        DistAcc<ELEM, METRIC> acc;
        acc.init();
        const bool mine = lane < cnt && (want || !masked);
        if (c.stage2 != nullptr) {
            // two chunk buffers: the next chunk's bulk copies are issued before this chunk is accumulated, so the memory
            // round trip of chunk i + 1 hides behind the arithmetic of chunk i (the build's searches are latency bound)
            const uint32_t m = __ballot_sync(FULL, mine);
            if (m) {
                const uint32_t n_mine = (uint32_t)__popc(m);
                const uint8_t *src = g.vectors + (size_t)(mine ? c.nbr[lane] : 0) * g.row_stride;
                auto issue = [&](uint32_t off, int b) {
                    const uint32_t cb = min(c.chunk_bytes, g.row_stride - off);
                    if (lane == 0) mbar_arrive_expect_tx(c.mbar + b, n_mine * cb);
                    __syncwarp();
                    if (mine) tma_row_g2s((b ? c.stage2 : c.stage) + (size_t)lane * c.chunk_stride, src + off, cb, c.mbar + b, c.policy);
                };
                issue(0, 0);
                int b = 0;
                for (uint32_t off = 0; off < g.row_stride; off += c.chunk_bytes, b ^= 1) {
                    const uint32_t cb = min(c.chunk_bytes, g.row_stride - off);
                    if (off + c.chunk_bytes < g.row_stride) issue(off + c.chunk_bytes, b ^ 1);
                    if (b == 0) { mbar_wait(c.mbar, c.phase); c.phase ^= 1u; }
                    else { mbar_wait(c.mbar + 1, c.phase2); c.phase2 ^= 1u; }
                    PH_TICK(c, 2);
                    if (mine) acc.add(c.qsm + off, (b ? c.stage2 : c.stage) + (size_t)lane * c.chunk_stride, (int)(cb / 16));
                    __syncwarp();
                    PH_TICK(c, 3);
                }
            }
            if (mine) d = acc.finish();
            return d;
        }
        for (uint32_t off = 0; off < g.row_stride; off += c.chunk_bytes) {
            const uint32_t cb = min(c.chunk_bytes, g.row_stride - off);
            gather_chunk(c, g, mine, off, cb);
            PH_TICK(c, 2);
            if (mine) acc.add(c.qsm + off, c.stage + (size_t)lane * c.chunk_stride, (int)(cb / 16));
            __syncwarp();
            PH_TICK(c, 3);
        }
        if (mine) d = acc.finish();
    }
    return d;
}

// search_layer (graph/mod.rs:161-255 == scan.rs:302-448 without the iterative-scan bookkeeping).
// Precondition: S[0, len) holds the entry points (unexpanded), their ids are in the visited set.
// FLAGGED: also honour GraphView::flags (an index a vacuum has touched); the default instantiation carries no trace of it.
// SKIP: the vacuum repair search (scan.rs:314-336,416-429 `skip_count`): members of the skip set -- elements left without
// heap TIDs and the element being repaired -- enter W and C like anyone else but do not count towards ef.  Since the
// reference evicts the furthest of W (whatever it is) only when a COUNTED member pushed w_len past ef, W is always the
// first min(counted pushes, ef) + (skip pushes) entries of the sorted list: the boundary c.ef simply moves out by one
// with every skip member that is pushed, and "W full" stays len >= ef.  One wrinkle: a layer seeded with MORE than ef
// counted entry points starts with w_len > ef (nothing is evicted while seeding), and then `if w_len > ef` fires after
// every push, skip members included, each time taking w_len down by one: the first `excess` skip pushes evict like
// counted ones and only the later ones move the boundary.
template <int ELEM, int METRIC, bool FLAGGED = false, bool SKIP = false, class LIST>
__device__ __forceinline__ void search_layer(WarpCtx &c, LIST &S, const GraphView &g, int layer, uint32_t vis_mask,
                                             uint32_t vis_limit) {
    const int lane = c.lane;
    const int lm = layer == 0 ? 2 * g.m : g.m;
    uint32_t vis_count = (uint32_t)c.len;
    // a lane that lost a same-slot race inside the warp (see below) still has to store its id; that is checked just
    // before the NEXT probe of the table, by which time the claim's round trip has long completed
    bool lost_fresh = false;
    uint32_t lost_old = VIS_EMPTY, lost_id = 0, lost_from = 0;
    while (true) {
        uint64_t ckey;
        PH_START(c);
        if (!S.pop_nearest(c, ckey)) break;   // every remaining candidate would fail `c.distance > f_dist`
        const uint32_t cid = key_id(ckey);
        if (c.rs_log) {
            // build: (expanded node, W's furthest distance at the start of the expansion or ~0 when W is not full, layer)
            if (lane == 0 && c.rs_n < c.rs_cap) {
                int32_t *e = c.rs_log + (size_t)c.rs_ent * (size_t)c.rs_n;
                e[0] = (int32_t)cid;
                e[1] = c.len >= c.ef ? (int32_t)c.f_ord : -1;
                e[2] = layer;
            }
            c.rs_n++;
        }
        const bool log_full = c.len >= c.ef;        // W full at the start of this expansion
        const uint32_t log_f = c.f_ord;
        const int32_t *row;
        if (layer == 0) {
            row = g.adj0 + (size_t)cid * (size_t)lm;
            c.n_exp0++;
        } else {
            row = g.adj_up + ((size_t)__ldg(g.upper_base + cid) + (size_t)(layer - 1)) * (size_t)lm;
            c.n_expu++;
        }
        for (int base = 0; base < lm; base += 32) {
            const int j = base + lane;
            const int32_t nb = j < lm ? __ldg(row + j) : -1;
            const bool valid = nb >= 0;                 // lists are dense prefixes (scan.rs:273-279)
            const uint32_t vm = __ballot_sync(FULL, valid);
            if (!vm) break;
            const int cnt = __popc(vm);                 // valid lanes are exactly 0 .. cnt-1
            PH_TICK(c, 0);
            // (1) visited test-and-set: the first probe is ISSUED here and only consumed after the distances,
            //     so its L2 round trip overlaps the row gather instead of preceding it.  Rows of neighbours
            //     that turn out to be visited already are fetched speculatively and ignored.
            if (lost_fresh && lost_old != VIS_EMPTY) vis_insert_from(c.vis, vis_mask, lost_from, lost_id);
            lost_fresh = false;
            const uint32_t h = vis_home((uint32_t)nb, vis_mask);
            uint4 bk = make_uint4(VIS_EMPTY, VIS_EMPTY, VIS_EMPTY, VIS_EMPTY);
            if (valid) bk = __ldcg(reinterpret_cast<const uint4 *>(c.vis + h));   // L2 (the table is written with atomics)
            int lvl = 255;
            if (layer > 0 && valid) lvl = __ldg(g.level + nb);
            uint32_t eflags = 0;                                          // only a vacuumed index carries flags
            if constexpr (FLAGGED) { if (g.flags != nullptr && valid) eflags = __ldg(g.flags + nb); }
            if (valid) c.nbr[lane] = nb;
            __syncwarp();
            // Two orders of the same steps.  Speculative (c.spec, the default; the build's log also needs every entry's
            // distance): gather all rows while the probes are in flight, resolve afterwards.  Probe-first
            // (PGVRX_NO_SPEC_GATHER): resolve the probes, then fetch only the rows that are evaluated -- 17 % less
            // L2 -> SM traffic, but measured 1 % slower at C2 (the kernel is latency bound, not fabric bound).
            float d = 0.0f;
            int nrow0 = 0;
            if (c.hybrid) {
                // (2h) the first round's rows are requested NOW, before the probes are consumed: the two round trips overlap
                nrow0 = min(c.R, cnt);
                if (lane == 0) mbar_arrive_expect_tx(c.mbar, (uint32_t)nrow0 * g.row_stride);
                __syncwarp();
                if (lane < nrow0)
                    tma_row_g2s(c.stage + (size_t)lane * c.stage_stride, g.vectors + (size_t)nb * g.row_stride, g.row_stride, c.mbar, c.policy);
            } else if (c.spec) {
                // (2) gather + one distance per lane (lane r <-> neighbour r of this chunk)
                d = eval_neighbours<ELEM, METRIC>(c, g, cnt, valid, false);
                // build log: which entries of this list are nearer than W's furthest distance was when the expansion
                // started (every valid lane has its distance, visited or not) -- the host replays against it
                if (c.rs_log && c.rs_n <= c.rs_cap) {
                    const uint32_t nm = __ballot_sync(FULL, valid && log_full && f32_ord(d) < log_f);
                    if (lane == 0) c.rs_log[(size_t)c.rs_ent * (size_t)(c.rs_n - 1) + 3 + (base >> 5)] = (int32_t)nm;
                }
            }
            // (3) resolve the visited probes (visited BEFORE the level test, graph/mod.rs:206-216)
            const uint32_t id = (uint32_t)nb;
            const bool present = bk.x == id || bk.y == id || bk.z == id || bk.w == id;
            const int free_slot = bk.x == VIS_EMPTY ? 0 : bk.y == VIS_EMPTY ? 1 : bk.z == VIS_EMPTY ? 2 : bk.w == VIS_EMPTY ? 3 : -1;
            bool fresh = valid && !present && free_slot >= 0;
            // claim the free slot; the result only matters when another lane of this warp claimed the same
            // slot in this round (different id), which is checked before the next probe of the table
            uint32_t claim_old = VIS_EMPTY;
            if (fresh) claim_old = atomicCAS(c.vis + h + free_slot, VIS_EMPTY, id);
            // home bucket full of other ids (rare at <= 40 % load): continue probing synchronously
            if (valid && !present && free_slot < 0) fresh = vis_insert_from(c.vis, vis_mask, (h + 4) & vis_mask, id);
            const uint32_t fm = __ballot_sync(FULL, fresh);
            vis_count += __popc(fm);
            c.n_vis += __popc(fm);
            if (vis_count > vis_limit) {
                if (c.hybrid) { mbar_wait(c.mbar, c.phase); c.phase ^= 1u; }      // round 0's rows are still landing
                c.status |= ST_VIS_OVERFLOW;
                return;
            }
            // graph/mod.rs:214, scan.rs:410; a deleted element fails load_element before its distance (scan.rs:179)
            const bool pass = fresh && lvl >= layer && !(eflags & ELEM_FLAG_DELETED);
            c.n_dist += __popc(__ballot_sync(FULL, pass));
            PH_TICK(c, 1);
            if (c.hybrid) {
                // (2h') round 0 has been in flight since before the probes; later rounds fetch only what is evaluated
                mbar_wait(c.mbar, c.phase);
                c.phase ^= 1u;
                PH_TICK(c, 2);
                if (pass && lane < nrow0) d = row_distance<ELEM, METRIC>(c.qsm, c.stage + (size_t)lane * c.stage_stride, c.nvec);
                __syncwarp();
                PH_TICK(c, 3);
                for (int s = nrow0; s < cnt; s += c.R) {
                    const int nrow = min(c.R, cnt - s);
                    gather_rows_masked(c, g, s, nrow, pass);
                    PH_TICK(c, 2);
                    if (pass && lane >= s && lane < s + nrow)
                        d = row_distance<ELEM, METRIC>(c.qsm, c.stage + (size_t)(lane - s) * c.stage_stride, c.nvec);
                    __syncwarp();
                    PH_TICK(c, 3);
                }
            } else if (!c.spec) {
                // (2') only the rows that are evaluated
                d = eval_neighbours<ELEM, METRIC>(c, g, cnt, pass, true);
            }
            // (4) accept in neighbour order (graph/mod.rs:226-242): the threshold only tightens and W only
            //     fills up, so lanes failing against the CURRENT threshold can never be accepted later
            const uint32_t dord = f32_ord(d);
            uint32_t am = __ballot_sync(FULL, pass && (c.len < c.ef || dord < c.f_ord));
            while (am) {
                const int r = __ffs(am) - 1;
                am &= am - 1;
                const float dr = __shfl_sync(FULL, d, r);
                const int32_t idr = __shfl_sync(FULL, nb, r);
                const bool always_add = c.len < c.ef;
                if (always_add || f32_ord(dr) < c.f_ord) {
                    if constexpr (SKIP) {
                        if (idr == c.self_id || (__ldg(g.flags + idr) & ELEM_FLAG_NO_HEAPTIDS)) {
                            if (c.excess > 0) c.excess -= 1; else c.ef += 1;
                        }
                        if (c.ef >= c.cap) { c.status |= ST_TIE_OVERFLOW; return; }   // more skip members than the list holds
                    }
                    S.insert(c, make_key(dr, (uint32_t)idr));
                }
            }
            // lost a same-slot race inside the warp: the id is new (decided above) but still has to be stored -- deferred
            lost_fresh = fresh && free_slot >= 0;
            lost_old = claim_old;
            lost_id = id;
            lost_from = (h + (uint32_t)free_slot + 1u) & vis_mask;
            PH_TICK(c, 4);
        }
    }
}

}  // namespace pgvrx
