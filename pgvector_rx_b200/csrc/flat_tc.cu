// flat_tc.cu -- K4b: candidate generation for the exact flat kNN of large fp32 L2 / inner-product batches on
// the 5th-generation tensor cores, as a 3xTF32 split GEMM:
//
//     q . x  ~=  q_hi . x_hi + q_hi . x_lo + q_lo . x_hi        (q = q_hi + q_lo + O(2^-22 |q|), same for x)
//
// with tf32 operands staged in shared memory by TMA (128-byte swizzle, K-major), tcgen05.mma (kind::tf32,
// M = 128 queries x N = 128 base rows per instruction group) accumulating in TMEM, and an epilogue that reads
// the accumulator back with tcgen05.ld, forms the approximate distance and keeps the KP nearest rows of each
// query.  The result is NOT the answer: the candidates are re-ranked with the reference's exact sequential
// arithmetic (rerank_kernel), and a query is only accepted when the error bound proves that no row outside
// its candidate list can belong to the top k (tc_certify_kernel); otherwise it is re-run on the exact path.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2-5 =
// epilogue (one TMEM lane quarter each; thread <-> query row).  Three pipelines: smem full/empty,
// TMEM full/empty (two 128-column accumulators), and a static persistent schedule over (query tile, base split).
#include <cuda.h>

#include "hnsw_device.cuh"
#include "hnsw_kernels.h"

namespace pgvrx {

constexpr int TC_TILE_BYTES = TC_M * TC_KB * 4;          // one 128 x 32 fp32 operand tile (16 KiB)
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;        // q_hi, q_lo, x_hi, x_lo
constexpr uint32_t TC_TMEM_COLS = 256;                   // two accumulators of TC_N fp32 columns
constexpr int TC_MAX_STAGES = 4;
constexpr int TC2_M = 256, TC2_N = 256;                  // 2-CTA tile: 128 query rows per CTA, 128 of the 256 x rows staged per CTA
constexpr int TC2_STAGES = 4;                            // ring slots of ONE operand half-tile (16 KiB) each
constexpr int TC_A_RES_BYTES = 4 * 2 * TC_TILE_BYTES;    // resident query tile: up to 4 K blocks of (hi | lo)

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait with back-off: the single-thread producer / MMA-issuer roles share their SM sub-partitions with epilogue warps,
// and a tight try_wait loop would take issue slots away from them
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, unsigned ns) {
    const uint32_t a = smem_u32(bar);
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile with 128-byte swizzle: rows are 128 B (32 tf32), 8-row groups are 1024 B apart
// (cute::UMMA::SmemDescriptor: start >> 4 | LBO 1 | SBO 64 << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61)
__device__ __forceinline__ uint64_t tc_smem_desc(const void *tile) {
    const uint32_t a = smem_u32(tile);
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10, K-major both, N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t tc_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- operand preparation: a -> (tf32 hi, tf32 lo) and |a|^2 -------------------------------------------------
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// one warp per row; rows beyond n_rows (padding up to a tile multiple) and columns beyond dim are zero
__global__ void __launch_bounds__(256) tc_split_kernel(const uint8_t *__restrict__ src, uint32_t src_stride, int64_t n_rows,
                                                       int64_t n_pad, int dim, int d_pad, float *__restrict__ hi,
                                                       float *__restrict__ lo, float *__restrict__ norm, uint32_t *norm_max,
                                                       uint32_t *hint_init) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_pad) return;
    const float *a = reinterpret_cast<const float *>(src + (size_t)row * src_stride);
    float s = 0.0f;
    for (int j = lane; j < d_pad; j += 32) {
        float v = (row < n_rows && j < dim) ? a[j] : 0.0f;
        const float h = to_tf32(v);
        const float l = to_tf32(__fsub_rn(v, h));
        hi[(size_t)row * d_pad + j] = h;
        lo[(size_t)row * d_pad + j] = l;
        s = __fmaf_rn(v, v, s);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) {
        norm[row] = s;
        if (norm_max) atomicMax(norm_max, __float_as_uint(s));     // s >= 0: float order == unsigned order
        if (hint_init && row < n_rows) hint_init[row] = 0xFF800000u;   // f32_ord(+inf): no threshold known yet
    }
}

// ---- epilogue of one unit (query tile x base split), shared by the 1-CTA and the 2-CTA kernels ------------------
// TN = accumulator width (base rows per tile).  Thread <-> query row for the arithmetic; the candidate set of a query
// is an UNORDERED set of TC_KP = 32 keys in shared memory, one per lane, maintained by the whole warp: a qualifying
// element replaces the set's maximum (two REDUX per 32-bit half), so an insertion is ~20 warp instructions with no
// divergent loop.  Qualifying elements are rare once the sets have warmed up (KP * ln(rows / KP) per query).
// KL > 0 is the hint pre-pass: no sets at all -- every lane keeps the KL smallest approximate distances of ITS query in
// registers (a branch-free insertion network, entered for a column only when some lane of the warp has a taker) and
// leaves the KL-th smallest in hint[q].  The set-maintaining form spent 79 warp instructions per insertion on cross-lane
// maxima; the sample pre-pass of a 125k-row shard took 0.39 ms for 0.04 ms of MMAs.
template <int TN, bool REMOTE_EMPTY, int KL>
__device__ __forceinline__ void tc_epilogue_unit(const TcParams &p, int q_tile0, int sp, int nt0, int nt1, uint32_t tmem_base,
                                                 float *sxn, uint64_t *Lw, uint64_t *tfull, uint64_t *tempty,
                                                 uint32_t tempty_leader_addr, int &buf, uint32_t &aphase, int quarter, int lane, int et) {
    const int r = quarter * 32 + lane;                              // query row inside the tile
    const int q = q_tile0 + r;
    const bool qvalid = q < p.nq;
    const float qn = qvalid ? p.qnorm[q] : 0.0f;
    uint32_t tau_hi = 0xFFFFFFFFu;                                   // this lane's query: distance bits of its set's maximum (all-ones = not full)
    float tau_f = __int_as_float(0x7f800000);                        // the same threshold as a distance
    // Thresholds travel between the units of a query: hint[q] is the smallest full-set maximum any unit has
    // reached so far.  Rejecting against it is sound because certification uses the minimum over all units'
    // final maxima, which is <= every value the hint ever took.
    float xn_cur[TN / 128];
#pragma unroll
    for (int h = 0; h < TN / 128; h++) xn_cur[h] = p.xnorm[(int64_t)nt0 * TN + h * 128 + et];
    if constexpr (KL > 0) {
        float top[KL];
#pragma unroll
        for (int i = 0; i < KL; i++) top[i] = __int_as_float(0x7f800000);
        for (int nt = nt0; nt < nt1; nt += p.tile_step) {
            const int64_t row0 = (int64_t)nt * TN;
#pragma unroll
            for (int h = 0; h < TN / 128; h++) {
                sxn[buf * TN + h * 128 + et] = xn_cur[h];
                if (nt + p.tile_step < nt1) xn_cur[h] = p.xnorm[row0 + (int64_t)p.tile_step * TN + h * 128 + et];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(tfull + buf, aphase);
            tc_fence_after();
            const int valid_cols = (int)min((int64_t)TN, p.n_rows - row0);
            for (int c = 0; c < TN; c += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TN + c), v);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float s = __uint_as_float(v[j]);
                    const float dist = p.metric == METRIC_L2SQ ? __fmaf_rn(-2.0f, s, qn + sxn[buf * TN + c + j]) : -s;
                    const bool take = qvalid && c + j < valid_cols && dist < top[KL - 1];
                    if (__any_sync(FULL, take)) {
                        float x = take ? dist : __int_as_float(0x7f800000);
#pragma unroll
                        for (int i = 0; i < KL; i++) {                       // sorted insert, the largest falls off the end
                            const float lo = fminf(top[i], x);
                            x = fmaxf(top[i], x);
                            top[i] = lo;
                        }
                    }
                }
            }
            tc_fence_before();
            if (REMOTE_EMPTY) {
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty_leader_addr + (uint32_t)buf * 8u) : "memory");
            } else {
                mbar_arrive(tempty + buf);
            }
            buf ^= 1;
            if (buf == 0) aphase ^= 1u;
        }
        // the KL-th smallest of the sample (if the sample held that many rows): rows are rejected against it from now on
        if (qvalid && top[KL - 1] < __int_as_float(0x7f800000)) atomicMin(p.hint + q, f32_ord(top[KL - 1]));
        __syncwarp();
        return;
    }
    int app_n = 0;                                                   // append mode: rows this lane's query has appended in this unit
    int32_t *app_seg = p.app_ids ? p.app_ids + ((size_t)(qvalid ? q : 0) * p.n_splits + sp) * p.app_cap : nullptr;
    // a set of kp_eff < 32 keys (the hint pre-pass): the other slots hold key 0, which nothing replaces and which is never
    // the maximum once a real key is present
#pragma unroll 4
    for (int i = 0; i < 32; i++) Lw[i * TC_KP + lane] = lane < p.kp_eff ? ~0ull : 0ull;
    __syncwarp();
    for (int nt = nt0; nt < nt1; nt += p.tile_step) {
        const int64_t row0 = (int64_t)nt * TN;
#pragma unroll
        for (int h = 0; h < TN / 128; h++) {
            sxn[buf * TN + h * 128 + et] = xn_cur[h];
            if (nt + p.tile_step < nt1) xn_cur[h] = p.xnorm[row0 + (int64_t)p.tile_step * TN + h * 128 + et];   // in flight while this tile is processed
        }
        const uint32_t hint_ord = qvalid ? __ldcg(p.hint + q) : 0xFF800000u;
        asm volatile("bar.sync 1, 128;" ::: "memory");              // the four epilogue warps only
        mbar_wait(tfull + buf, aphase);
        tc_fence_after();
        const float hint_f = ord_f32(hint_ord);
        float thr_f = fminf(tau_f, hint_f);
        const int valid_cols = (int)min((int64_t)TN, p.n_rows - row0);   // rows beyond n_rows are padding
        for (int c = 0; c < TN; c += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TN + c), v);
            if (p.dbg_scores) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if (qvalid && c + j < valid_cols) {
                        const float s = __uint_as_float(v[j]);
                        p.dbg_scores[(size_t)q * p.n_rows + row0 + c + j] =
                            p.metric == METRIC_L2SQ ? __fmaf_rn(-2.0f, s, qn + sxn[buf * TN + c + j]) : -s;
                    }
                }
            }
            // fast path: one fused multiply-add and one compare per element, no cross-lane traffic
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const float s = __uint_as_float(v[j]);
                const float dist = p.metric == METRIC_L2SQ ? __fmaf_rn(-2.0f, s, qn + sxn[buf * TN + c + j]) : -s;
                mask |= (dist <= thr_f ? 1u : 0u) << j;             // superset of "nearer than the set's maximum" (decided below)
            }
            const int nvalid = valid_cols - c;
            if (!qvalid || nvalid <= 0) mask = 0;
            else if (nvalid < 32) mask &= (1u << nvalid) - 1u;
            if (p.app_ids != nullptr) {
                // append mode: the threshold is the pre-pass hint and does not move; a lane appends its own query's rows to
                // the segment this unit owns in the query's list -- a private counter, plain stores, no cross-lane work and
                // no atomics (a returning atomic per append would stall the warp for an L2 round trip each time)
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (app_n < p.app_cap) app_seg[app_n] = (int32_t)((uint32_t)(row0 + c + j) + (uint32_t)p.id_base);
                    app_n++;
                }
                continue;
            }
            if (__ballot_sync(FULL, mask != 0)) {
                // slow path (rare once the sets have warmed up): column by column over the union of the masks
                uint32_t um = __reduce_or_sync(FULL, mask);
                while (um) {
                    const int j = __ffs(um) - 1;
                    um &= um - 1;
                    uint32_t sv;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(sv)
                                 : "r"(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TN + c + j)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float s = __uint_as_float(sv);
                    const float dist = p.metric == METRIC_L2SQ ? __fmaf_rn(-2.0f, s, qn + sxn[buf * TN + c + j]) : -s;
                    const uint32_t khi = f32_ord(dist), klo = (uint32_t)(row0 + c + j);
                    // Sets are ordered by the distance bits alone (an element that ties with the set's maximum is not
                    // taken: it is >= the bound either way), which halves the REDUX chain of an insertion; and two
                    // queries' insertions run interleaved -- their sets are independent -- so that the chains overlap.
                    uint32_t am = __ballot_sync(FULL, ((mask >> j) & 1u) && khi < tau_hi);
                    while (am) {
                        const int s0 = __ffs(am) - 1;
                        am &= am - 1;
                        const bool two = am != 0;
                        const int s1 = two ? __ffs(am) - 1 : s0;
                        if (two) am &= am - 1;
                        const uint32_t nhi0 = __shfl_sync(FULL, khi, s0), nlo0 = __shfl_sync(FULL, klo, s0);
                        const uint32_t nhi1 = __shfl_sync(FULL, khi, s1), nlo1 = __shfl_sync(FULL, klo, s1);
                        uint32_t chi0 = (uint32_t)(Lw[s0 * TC_KP + lane] >> 32);
                        uint32_t chi1 = (uint32_t)(Lw[s1 * TC_KP + lane] >> 32);
                        // the set's maximum is replaced by the new key
                        const uint32_t m0 = __reduce_max_sync(FULL, chi0);
                        const uint32_t m1 = __reduce_max_sync(FULL, chi1);
                        const int o0 = __ffs(__ballot_sync(FULL, chi0 == m0)) - 1;
                        const int o1 = __ffs(__ballot_sync(FULL, chi1 == m1)) - 1;
                        if (lane == o0) { chi0 = nhi0; Lw[s0 * TC_KP + lane] = ((uint64_t)nhi0 << 32) | nlo0; }
                        if (two && lane == o1) { chi1 = nhi1; Lw[s1 * TC_KP + lane] = ((uint64_t)nhi1 << 32) | nlo1; }
                        const uint32_t t0 = __reduce_max_sync(FULL, chi0);
                        const uint32_t t1 = __reduce_max_sync(FULL, chi1);
                        if (lane == s0 || (two && lane == s1)) {
                            tau_hi = lane == s0 ? t0 : t1;
                            if (tau_hi != 0xFFFFFFFFu) {                      // the set is full: its maximum is a threshold
                                tau_f = ord_f32(tau_hi);
                                thr_f = fminf(thr_f, tau_f);
                                atomicMin(p.hint + q, tau_hi);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        tc_fence_before();
        if (REMOTE_EMPTY) {
            // 2-CTA: the MMA issuer lives in the leader CTA and waits for the epilogues of both CTAs
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty_leader_addr + (uint32_t)buf * 8u) : "memory");
        } else {
            mbar_arrive(tempty + buf);
        }
        buf ^= 1;
        if (buf == 0) aphase ^= 1u;
    }
    __syncwarp();
    // candidates of query (quarter * 32 + i), one per lane; -1 where the split had fewer than KP rows
    // (the hint pre-pass has no outputs but p.hint)
    if (p.cand_ids) {
        for (int i = 0; i < 32; i++) {
            const int qi = q_tile0 + quarter * 32 + i;
            if (qi >= p.nq) break;
            const uint64_t cur = Lw[i * TC_KP + lane];
            p.cand_ids[((size_t)qi * p.n_splits + sp) * TC_KP + lane] = cur == ~0ull ? -1 : (int32_t)((uint32_t)cur + (uint32_t)p.id_base);
        }
    }
    if (qvalid && p.app_cnt) p.app_cnt[(size_t)q * p.n_splits + sp] = app_n;
    // every row of this split that was not kept has an approximate distance >= this bound
    if (qvalid && p.cand_bound) p.cand_bound[(size_t)sp * p.nq + q] = tau_hi == 0xFFFFFFFFu ? __int_as_float(0x7f800000) : ord_f32(tau_hi);
    __syncwarp();
}

// ---- the GEMM + top-KP kernel -------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
flat_tc_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
               const __grid_constant__ CUtensorMap map_xhi, const __grid_constant__ CUtensorMap map_xlo, TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128-byte swizzle atoms
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *stages = smem;                                                   // TC_STAGES x (q_hi | q_lo | x_hi | x_lo)
    float *sxn = reinterpret_cast<float *>(smem + TC_STAGES * TC_STAGE_BYTES); // 2 x TC_N row norms
    uint64_t *lists = reinterpret_cast<uint64_t *>(sxn + 2 * TC_N);           // 128 x KP keys
    uint64_t *bars = lists + (size_t)TC_M * p.kp;
    uint64_t *full = bars, *empty = bars + TC_MAX_STAGES, *tfull = bars + 2 * TC_MAX_STAGES, *tempty = tfull + 2;
    uint64_t *afull = tempty + 2, *aempty = afull + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(aempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_MAX_STAGES; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, TC_M); }
        mbar_init(afull, 1);
        mbar_init(aempty, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_units = p.n_mtiles * p.n_splits;
    const int n_kb = p.d_pad / TC_KB;

    // Operand staging.  When the whole K extent of a query tile fits (d_pad <= 128) the tile's q_hi / q_lo blocks
    // stay resident for the unit and only x_hi / x_lo stream through the ring (half the shared-memory fill traffic);
    // otherwise all four operand tiles of a K block share a ring slot.
    const bool a_res = p.a_resident != 0;
    const int n_stages = a_res ? (TC_STAGES * TC_STAGE_BYTES - TC_A_RES_BYTES) / (2 * TC_TILE_BYTES) : TC_STAGES;
    const uint32_t stage_bytes = a_res ? 2u * TC_TILE_BYTES : (uint32_t)TC_STAGE_BYTES;
    uint8_t *ring = a_res ? stages + TC_A_RES_BYTES : stages;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint64_t pol_x = l2_policy_evict_first(), pol_q = l2_policy_evict_last();
            int stage = 0;
            uint32_t phase = 0, aph = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int mt = u % p.n_mtiles, sp = u / p.n_mtiles;
                const int nt0 = sp * p.ntiles_per_split;
                const int nt1 = min(nt0 + p.ntiles_per_split, p.n_ntiles);
                if (a_res) {
                    mbar_wait(aempty, aph ^ 1u);                    // the MMAs of the previous unit are done with the query tile
                    mbar_arrive_expect_tx(afull, (uint32_t)n_kb * 2u * TC_TILE_BYTES);
                    for (int kb = 0; kb < n_kb; kb++) {
                        tma_load_2d(stages + (size_t)kb * 2 * TC_TILE_BYTES, &map_qhi, kb * TC_KB, mt * TC_M, afull, pol_q);
                        tma_load_2d(stages + (size_t)kb * 2 * TC_TILE_BYTES + TC_TILE_BYTES, &map_qlo, kb * TC_KB, mt * TC_M, afull, pol_q);
                    }
                    aph ^= 1u;
                }
                for (int nt = nt0; nt < nt1; nt += p.tile_step) {
                    for (int kb = 0; kb < n_kb; kb++) {
                        mbar_wait(empty + stage, phase ^ 1u);
                        uint8_t *st = ring + (size_t)stage * stage_bytes;
                        mbar_arrive_expect_tx(full + stage, stage_bytes);
                        if (!a_res) {
                            tma_load_2d(st, &map_qhi, kb * TC_KB, mt * TC_M, full + stage, pol_q);
                            tma_load_2d(st + TC_TILE_BYTES, &map_qlo, kb * TC_KB, mt * TC_M, full + stage, pol_q);
                            st += 2 * TC_TILE_BYTES;
                        }
                        tma_load_2d(st, &map_xhi, kb * TC_KB, nt * TC_N, full + stage, pol_x);
                        tma_load_2d(st + TC_TILE_BYTES, &map_xlo, kb * TC_KB, nt * TC_N, full + stage, pol_x);
                        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = tc_idesc();
            int stage = 0, buf = 0;
            uint32_t phase = 0, aphase = 0, aph = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int sp = u / p.n_mtiles;
                const int nt0 = sp * p.ntiles_per_split;
                const int nt1 = min(nt0 + p.ntiles_per_split, p.n_ntiles);
                if (a_res) {
                    mbar_wait(afull, aph);
                    tc_fence_after();
                    aph ^= 1u;
                }
                for (int nt = nt0; nt < nt1; nt += p.tile_step) {
                    mbar_wait(tempty + buf, aphase ^ 1u);          // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * TC_N;
                    for (int kb = 0; kb < n_kb; kb++) {
                        mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint8_t *st = ring + (size_t)stage * stage_bytes;
                        const uint8_t *qa = a_res ? stages + (size_t)kb * 2 * TC_TILE_BYTES : st;
                        const uint8_t *xb = a_res ? st : st + 2 * TC_TILE_BYTES;
                        const uint64_t qhi = tc_smem_desc(qa), qlo = tc_smem_desc(qa + TC_TILE_BYTES);
                        const uint64_t xhi = tc_smem_desc(xb), xlo = tc_smem_desc(xb + TC_TILE_BYTES);
                        // small cross terms first, then the main product (8 tf32 = 32 bytes = +2 in the descriptor per K step)
#pragma unroll
                        for (int k = 0; k < TC_KB / 8; k++) tc_mma_tf32(d_tmem, qlo + 2 * k, xhi + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
                        for (int k = 0; k < TC_KB / 8; k++) tc_mma_tf32(d_tmem, qhi + 2 * k, xlo + 2 * k, idesc, 1u);
#pragma unroll
                        for (int k = 0; k < TC_KB / 8; k++) tc_mma_tf32(d_tmem, qhi + 2 * k, xhi + 2 * k, idesc, 1u);
                        tc_commit(empty + stage);                   // frees the ring slot when these MMAs retire
                        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit(tfull + buf);                         // accumulator complete
                    buf ^= 1;
                    if (buf == 0) aphase ^= 1u;
                }
                if (a_res) tc_commit(aempty);                       // the query tile may be overwritten once all of this retired
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> approximate distance -> per-query top-KP (tc_epilogue_unit) =====
        const int quarter = warp & 3;                               // TMEM lanes [32 * quarter, +32) belong to this warp
        const int et = (warp - 2) * 32 + lane;                      // 0..127 among the epilogue threads
        uint64_t *Lw = lists + (size_t)(quarter * 32) * TC_KP;      // this warp's 32 sets: set i = Lw[i * 32 + lane]
        int buf = 0;
        uint32_t aphase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int mt = u % p.n_mtiles, sp = u / p.n_mtiles;
            const int nt0 = sp * p.ntiles_per_split;
            const int nt1 = min(nt0 + p.ntiles_per_split, p.n_ntiles);
            if (p.local_k == 8) tc_epilogue_unit<TC_N, false, 8>(p, mt * TC_M, sp, nt0, nt1, tmem_base, sxn, Lw, tfull, tempty, 0u, buf, aphase, quarter, lane, et);
            else if (p.local_k == 16) tc_epilogue_unit<TC_N, false, 16>(p, mt * TC_M, sp, nt0, nt1, tmem_base, sxn, Lw, tfull, tempty, 0u, buf, aphase, quarter, lane, et);
            else tc_epilogue_unit<TC_N, false, 0>(p, mt * TC_M, sp, nt0, nt1, tmem_base, sxn, Lw, tfull, tempty, 0u, buf, aphase, quarter, lane, et);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---- the same pipeline on CTA pairs: tcgen05.mma.cta_group::2 ------------------------------------------------------
// Two CTAs of a cluster (one SM pair) work on 256 queries x 256 base rows per tile.  Each CTA keeps ITS 128 queries'
// hi/lo K blocks resident and streams ITS half (128 rows) of every x tile; the leader's MMA instructions read both
// halves, so every SM fills its shared memory with half the x bytes per unit of tensor work -- the x stream is what
// bounds the 1-CTA kernel (L2 -> SM ~6 TB/s).  Each CTA's TMEM holds its 128 query rows x 256 columns, twice.
// Barriers: TMA completions of both CTAs land on the LEADER's full / afull barriers; tcgen05.commit is multicast to
// both CTAs' empty / aempty / tfull barriers; the epilogue threads of both CTAs arrive on the leader's tempty.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, int c0, int c1, uint32_t leader_bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
flat_tc2_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
                const __grid_constant__ CUtensorMap map_xhi, const __grid_constant__ CUtensorMap map_xlo, TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *a_res = smem;                                                     // 4 K blocks x (q_hi | q_lo) of this CTA's 128 queries
    uint8_t *ring = smem + TC_A_RES_BYTES;                                     // TC2_STAGES slots: x_hi half, x_lo half, x_hi half, ...
    float *sxn = reinterpret_cast<float *>(ring + TC2_STAGES * TC_TILE_BYTES);       // 2 x TC2_N row norms
    uint64_t *lists = reinterpret_cast<uint64_t *>(sxn + 2 * TC2_N);
    uint64_t *bars = lists + (size_t)TC_M * TC_KP;
    uint64_t *full = bars, *empty = bars + TC2_STAGES, *tfull = empty + TC2_STAGES, *tempty = tfull + 2;
    uint64_t *afull = tempty + 2, *aempty = afull + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(aempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if ((smem_u32(smem) & 1023u) != 0) __trap();                               // swizzle atoms need 1024-byte alignment
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC2_STAGES; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 2 * TC_M); }
        mbar_init(afull, 1);
        mbar_init(aempty, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                        // both CTAs' barriers exist before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_clusters = gridDim.x / 2, cid = blockIdx.x / 2;
    const int n_units = p.n_mtiles * p.n_splits;                               // n_mtiles counts 256-query tiles here
    const int n_kb = p.d_pad / TC_KB;
    const uint32_t stage_bytes = TC_TILE_BYTES;                                // fine-grained slots keep the x stream ahead of the MMAs

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own queries, own half of every x tile; completions go to the leader =====
        if (lane == 0) {
            const uint64_t pol_x = l2_policy_evict_first(), pol_q = l2_policy_evict_last();
            const uint32_t l_afull = map_to_cta(smem_u32(afull), 0);
            int stage = 0;
            uint32_t phase = 0, aph = 0;
            for (int u = cid; u < n_units; u += n_clusters) {
                const int mt = u % p.n_mtiles, sp = u / p.n_mtiles;
                const int nt0 = sp * p.ntiles_per_split;
                const int nt1 = min(nt0 + p.ntiles_per_split, p.n_ntiles);
                mbar_wait_backoff(aempty, aph ^ 1u, 200);
                if (leader) mbar_arrive_expect_tx(afull, 2u * (uint32_t)n_kb * 2u * TC_TILE_BYTES);
                for (int kb = 0; kb < n_kb; kb++) {
                    tma_load_2d_2sm(a_res + (size_t)kb * 2 * TC_TILE_BYTES, &map_qhi, kb * TC_KB, mt * TC2_M + (int)rank * TC_M, l_afull, pol_q);
                    tma_load_2d_2sm(a_res + (size_t)kb * 2 * TC_TILE_BYTES + TC_TILE_BYTES, &map_qlo, kb * TC_KB, mt * TC2_M + (int)rank * TC_M, l_afull, pol_q);
                }
                aph ^= 1u;
                for (int nt = nt0; nt < nt1; nt += p.tile_step) {
                    for (int kb = 0; kb < n_kb; kb++) {
#pragma unroll
                        for (int half = 0; half < 2; half++) {      // x_hi slot, then x_lo slot
                            mbar_wait_backoff(empty + stage, phase ^ 1u, 100);
                            if (leader) mbar_arrive_expect_tx(full + stage, 2u * stage_bytes);   // this CTA's half and the peer's
                            tma_load_2d_2sm(ring + (size_t)stage * stage_bytes, half == 0 ? &map_xhi : &map_xlo, kb * TC_KB,
                                            nt * TC2_N + (int)rank * TC_M, map_to_cta(smem_u32(full + stage), 0), pol_x);
                            if (++stage == TC2_STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && lane == 0) {
            // cute::UMMA::InstrDescriptor: F32 accumulate, TF32 x TF32, K-major, N = 256, M = 256 (128 rows per CTA)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC2_N >> 3) << 17) | ((uint32_t)(TC2_M >> 4) << 24);
            int stage = 0, buf = 0;
            uint32_t phase = 0, aphase = 0, aph = 0;
            for (int u = cid; u < n_units; u += n_clusters) {
                const int sp = u / p.n_mtiles;
                const int nt0 = sp * p.ntiles_per_split;
                const int nt1 = min(nt0 + p.ntiles_per_split, p.n_ntiles);
                mbar_wait_backoff(afull, aph, 100);
                tc_fence_after();
                aph ^= 1u;
                for (int nt = nt0; nt < nt1; nt += p.tile_step) {
                    mbar_wait_backoff(tempty + buf, aphase ^ 1u, 50);          // both CTAs' epilogues have drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * TC2_N;
                    for (int kb = 0; kb < n_kb; kb++) {
                        const uint8_t *qa = a_res + (size_t)kb * 2 * TC_TILE_BYTES;
                        const uint64_t qhi = tc_smem_desc(qa), qlo = tc_smem_desc(qa + TC_TILE_BYTES);
                        // x_hi slot: the cross term q_lo . x_hi, then the main product
                        mbar_wait_backoff(full + stage, phase, 50);
                        tc_fence_after();
                        const uint64_t xhi = tc_smem_desc(ring + (size_t)stage * stage_bytes);
#pragma unroll
                        for (int k = 0; k < TC_KB / 8; k++) tc_mma_tf32_2sm(d_tmem, qlo + 2 * k, xhi + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
                        for (int k = 0; k < TC_KB / 8; k++) tc_mma_tf32_2sm(d_tmem, qhi + 2 * k, xhi + 2 * k, idesc, 1u);
                        tc_commit_2sm(empty + stage);               // both CTAs' slots are free when these MMAs retire
                        if (++stage == TC2_STAGES) { stage = 0; phase ^= 1u; }
                        // x_lo slot: the cross term q_hi . x_lo
                        mbar_wait_backoff(full + stage, phase, 50);
                        tc_fence_after();
                        const uint64_t xlo = tc_smem_desc(ring + (size_t)stage * stage_bytes);
#pragma unroll
                        for (int k = 0; k < TC_KB / 8; k++) tc_mma_tf32_2sm(d_tmem, qhi + 2 * k, xlo + 2 * k, idesc, 1u);
                        tc_commit_2sm(empty + stage);
                        if (++stage == TC2_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit_2sm(tfull + buf);                     // both CTAs' accumulators are complete
                    buf ^= 1;
                    if (buf == 0) aphase ^= 1u;
                }
                tc_commit_2sm(aempty);
            }
        }
    } else {
        // ===== epilogue (both CTAs, each on its own 128 query rows) =====
        const int quarter = warp & 3;
        const int et = (warp - 2) * 32 + lane;
        uint64_t *Lw = lists + (size_t)(quarter * 32) * TC_KP;
        const uint32_t l_tempty = map_to_cta(smem_u32(tempty), 0);
        int buf = 0;
        uint32_t aphase = 0;
        for (int u = cid; u < n_units; u += n_clusters) {
            const int mt = u % p.n_mtiles, sp = u / p.n_mtiles;
            const int nt0 = sp * p.ntiles_per_split;
            const int nt1 = min(nt0 + p.ntiles_per_split, p.n_ntiles);
            if (p.local_k == 8)
                tc_epilogue_unit<TC2_N, true, 8>(p, mt * TC2_M + (int)rank * TC_M, sp, nt0, nt1, tmem_base, sxn, Lw, tfull, tempty, l_tempty, buf,
                                                 aphase, quarter, lane, et);
            else if (p.local_k == 16)
                tc_epilogue_unit<TC2_N, true, 16>(p, mt * TC2_M + (int)rank * TC_M, sp, nt0, nt1, tmem_base, sxn, Lw, tfull, tempty, l_tempty, buf,
                                                  aphase, quarter, lane, et);
            else
                tc_epilogue_unit<TC2_N, true, 0>(p, mt * TC2_M + (int)rank * TC_M, sp, nt0, nt1, tmem_base, sxn, Lw, tfull, tempty, l_tempty, buf,
                                                 aphase, quarter, lane, et);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                        // nobody leaves while the peer may still signal or read
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- certification ---------------------------------------------------------------------------------------------
// A query's exact top-k (from the re-rank of its candidates) is final when every row outside the candidate lists
// is provably further away than its k-th result: exact(r) >= approx(r) - eps >= min_s bound_s - eps > E_k.
__global__ void tc_certify_kernel(const float *__restrict__ topk_dists, const int32_t *__restrict__ topk_ids, int k,
                                  const float *__restrict__ cand_bound, int n_splits, int nq, const float *__restrict__ qnorm,
                                  const uint32_t *__restrict__ xnorm_max, float eps_rel, int metric, const uint32_t *__restrict__ hint,
                                  const int *__restrict__ app_cnt, int app_cap, int app_splits, int32_t *uncert_list, int *n_uncert) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    // Rows were rejected against min(the unit's own full-set maximum, hint[q] at the time); the hint only ever decreases
    // (atomicMin of full-set maxima, seeded by the sample pre-pass), so its final value is a lower bound of every
    // threshold used: every unkept row has an approximate distance >= min(split bounds, final hint).
    float bound = hint ? ord_f32(hint[q]) : __int_as_float(0x7f800000);
    for (int s = 0; s < n_splits; s++) bound = fminf(bound, cand_bound[(size_t)s * nq + q]);
    bool ok = true;
    bool dropped = false;                                              // rows under the threshold that found no room
    if (app_cnt != nullptr)
        for (int s = 0; s < app_splits; s++) dropped = dropped || app_cnt[(size_t)q * app_splits + s] > app_cap;
    if (dropped) ok = false;
    else if (bound < __int_as_float(0x7f800000)) {
        const float xm = __uint_as_float(*xnorm_max);
        const float eps = eps_rel * (qnorm[q] + xm);
        // the k-th result must exist and lie strictly inside the certified radius
        ok = topk_ids[(size_t)q * k + (k - 1)] >= 0 && topk_dists[(size_t)q * k + (k - 1)] < bound - eps;
    }
    if (!ok) uncert_list[atomicAdd(n_uncert, 1)] = q;
    (void)metric;
}

// gather / scatter of the few queries that have to be re-run on the exact path
__global__ void tc_gather_queries_kernel(const uint8_t *__restrict__ q, uint32_t q_stride, const int32_t *__restrict__ list, int n,
                                         uint8_t *__restrict__ out) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const uint8_t *src = q + (size_t)list[i] * q_stride;
    for (uint32_t b = threadIdx.x; b < q_stride; b += blockDim.x) out[(size_t)i * q_stride + b] = src[b];
}
__global__ void tc_scatter_results_kernel(const int32_t *__restrict__ ids, const float *__restrict__ dists, int k,
                                          const int32_t *__restrict__ list, int n, int32_t *out_ids, float *out_dists) {
    const int i = blockIdx.x;
    if (i >= n) return;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        out_ids[(size_t)list[i] * k + j] = ids[(size_t)i * k + j];
        out_dists[(size_t)list[i] * k + j] = dists[(size_t)i * k + j];
    }
}

// ---- launchers ---------------------------------------------------------------------------------------------------
size_t tc_smem_bytes(int kp) {
    return 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + 2 * TC_N * 4 + (size_t)TC_M * kp * 8 + (2 * TC_MAX_STAGES + 6) * 8 + 16;
}

cudaError_t launch_tc_split(const uint8_t *src, uint32_t src_stride, int64_t n_rows, int64_t n_pad, int dim, int d_pad, float *hi,
                            float *lo, float *norm, uint32_t *norm_max, uint32_t *hint_init, cudaStream_t stream) {
    if (n_pad <= 0) return cudaSuccess;
    tc_split_kernel<<<(unsigned)((n_pad + 7) / 8), 256, 0, stream>>>(src, src_stride, n_rows, n_pad, dim, d_pad, hi, lo, norm, norm_max,
                                                                     hint_init);
    return cudaGetLastError();
}

cudaError_t launch_flat_tc(const void *map_qhi, const void *map_qlo, const void *map_xhi, const void *map_xlo, const TcParams &p,
                           int grid, cudaStream_t stream) {
    const size_t smem = tc_smem_bytes(p.kp);
    cudaError_t e = cudaFuncSetAttribute(flat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    flat_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(*reinterpret_cast<const CUtensorMap *>(map_qhi),
                                                       *reinterpret_cast<const CUtensorMap *>(map_qlo),
                                                       *reinterpret_cast<const CUtensorMap *>(map_xhi),
                                                       *reinterpret_cast<const CUtensorMap *>(map_xlo), p);
    return cudaGetLastError();
}

size_t tc2_smem_bytes() {
    return (size_t)TC_A_RES_BYTES + (size_t)TC2_STAGES * TC_TILE_BYTES + 2 * TC2_N * 4 + (size_t)TC_M * TC_KP * 8 + (2 * TC2_STAGES + 6) * 8 + 16;
}

cudaError_t launch_flat_tc2(const void *map_qhi, const void *map_qlo, const void *map_xhi, const void *map_xlo, const TcParams &p,
                            int grid, cudaStream_t stream) {
    const size_t smem = tc2_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(flat_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    flat_tc2_kernel<<<grid, TC_THREADS, smem, stream>>>(*reinterpret_cast<const CUtensorMap *>(map_qhi),
                                                        *reinterpret_cast<const CUtensorMap *>(map_qlo),
                                                        *reinterpret_cast<const CUtensorMap *>(map_xhi),
                                                        *reinterpret_cast<const CUtensorMap *>(map_xlo), p);
    return cudaGetLastError();
}

cudaError_t launch_tc_certify(const float *topk_dists, const int32_t *topk_ids, int k, const float *cand_bound, int n_splits, int nq,
                              const float *qnorm, const uint32_t *xnorm_max, float eps_rel, int metric, const uint32_t *hint,
                              const int *app_cnt, int app_cap, int app_splits, int32_t *uncert_list, int *n_uncert, cudaStream_t stream) {
    tc_certify_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(topk_dists, topk_ids, k, cand_bound, n_splits, nq, qnorm, xnorm_max,
                                                            eps_rel, metric, hint, app_cnt, app_cap, app_splits, uncert_list, n_uncert);
    return cudaGetLastError();
}

cudaError_t launch_tc_gather_queries(const uint8_t *q, uint32_t q_stride, const int32_t *list, int n, uint8_t *out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tc_gather_queries_kernel<<<n, 128, 0, stream>>>(q, q_stride, list, n, out);
    return cudaGetLastError();
}

cudaError_t launch_tc_scatter_results(const int32_t *ids, const float *dists, int k, const int32_t *list, int n, int32_t *out_ids,
                                      float *out_dists, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tc_scatter_results_kernel<<<n, 32, 0, stream>>>(ids, dists, k, list, n, out_ids, out_dists);
    return cudaGetLastError();
}

}  // namespace pgvrx
