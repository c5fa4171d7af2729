// flat_knn.cu -- K4: exact kNN over the vector table (the operator-ordered sequential scan the
// reference's recall tests use as ground truth, SURVEY a20: tests/t/012...pl:83, 013...pl:95-98),
// the exact re-rank of candidate lists, and K6: the k-way top-k merge used across base-set shards.
//
// Distances are the reference's arithmetic bit for bit: every (query, row) pair is one strictly
// sequential f32 accumulation chain with separate multiply and add (vector.rs:518-567,
// halfvec.rs:687-733), integer popcounts for bit vectors (bitvec.rs:97-132).  The parallelism is
// across pairs, never inside one: a thread owns one base row of the tile and FLAT_QT queries, so it
// runs FLAT_QT independent chains.  Results are ordered by (distance, element id), like the oracle's
// sort of (operator value, ctid order).
//
// Roofline: FP32-issue bound, not HBM bound -- 2 packed + 1 scalar instruction per 2 elements per pair
// against 16 bytes of row data shared by FLAT_QT queries.
#include "hnsw_device.cuh"
#include "hnsw_kernels.h"

namespace pgvrx {

constexpr uint64_t FKEY_MAX = ~0ull;

// flat keys: ordered distance bits << 32 | id (no flag bit: ids up to 2^32 - 2)
__device__ __forceinline__ uint64_t flat_key(float d, uint32_t id) { return ((uint64_t)f32_ord(d) << 32) | id; }

// warp-cooperative insert of `key` into the ascending list L[0, len) of capacity k (smem); keeps the k smallest
__device__ __forceinline__ void topk_insert(uint64_t *L, int &len, int k, uint64_t key, int lane) {
    if (len == k && key >= L[k - 1]) return;
    int pos = 0;
    for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        const bool lt = i < len && L[i] < key;
        pos += __popc(__ballot_sync(FULL, lt));
    }
    const int newlen = len < k ? len + 1 : k;
    // move [pos, newlen - 1) up by one, highest chunk first
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

// ---------------------------------------------------------------------------------------------------
// K4: per (query tile, base split) partial top-k
// ---------------------------------------------------------------------------------------------------
template <int ELEM> struct FlatAcc { typedef float type; };
template <> struct FlatAcc<ELEM_BIT> { typedef uint32_t type; };

// one 16-byte vector of the thread's row against the same 16 bytes of FLAT_QT queries.
// qs is the transposed query tile: 16-byte vector v of query q lives at qs[(v * FLAT_QT + q)].
template <int ELEM, int METRIC>
__device__ __forceinline__ void flat_step(typename FlatAcc<ELEM>::type (&acc)[FLAT_QT], uint32_t &bb, const uint4 x,
                                          const uint4 *__restrict__ qs, int j) {
    if (ELEM == ELEM_F32) {
        const float2 x01 = make_float2(__uint_as_float(x.x), __uint_as_float(x.y));
        const float2 x23 = make_float2(__uint_as_float(x.z), __uint_as_float(x.w));
        const uint4 *qv = qs + (size_t)j * FLAT_QT;
#pragma unroll
        for (int q = 0; q < FLAT_QT; q++) {
            const uint4 a = qv[q];
            float v = reinterpret_cast<float &>(acc[q]);
            v = acc_step2<METRIC>(v, make_float2(__uint_as_float(a.x), __uint_as_float(a.y)), x01);
            v = acc_step2<METRIC>(v, make_float2(__uint_as_float(a.z), __uint_as_float(a.w)), x23);
            reinterpret_cast<float &>(acc[q]) = v;
        }
    } else if (ELEM == ELEM_F16) {
        // the row vector holds 8 halves; the query tile was widened to f32 when it was staged (exact)
        const float2 x01 = __half22float2(*reinterpret_cast<const __half2 *>(&x.x));
        const float2 x23 = __half22float2(*reinterpret_cast<const __half2 *>(&x.y));
        const float2 x45 = __half22float2(*reinterpret_cast<const __half2 *>(&x.z));
        const float2 x67 = __half22float2(*reinterpret_cast<const __half2 *>(&x.w));
        const uint4 *qv = qs + (size_t)(2 * j) * FLAT_QT;
#pragma unroll
        for (int q = 0; q < FLAT_QT; q++) {
            const uint4 a = qv[q], b = qv[FLAT_QT + q];
            float v = reinterpret_cast<float &>(acc[q]);
            v = acc_step2<METRIC>(v, make_float2(__uint_as_float(a.x), __uint_as_float(a.y)), x01);
            v = acc_step2<METRIC>(v, make_float2(__uint_as_float(a.z), __uint_as_float(a.w)), x23);
            v = acc_step2<METRIC>(v, make_float2(__uint_as_float(b.x), __uint_as_float(b.y)), x45);
            v = acc_step2<METRIC>(v, make_float2(__uint_as_float(b.z), __uint_as_float(b.w)), x67);
            reinterpret_cast<float &>(acc[q]) = v;
        }
    } else {
        const uint4 *qv = qs + (size_t)j * FLAT_QT;
        if (METRIC == METRIC_JACCARD) bb += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
#pragma unroll
        for (int q = 0; q < FLAT_QT; q++) {
            const uint4 a = qv[q];
            uint32_t v = reinterpret_cast<uint32_t &>(acc[q]);
            if (METRIC == METRIC_HAMMING) v += __popc(a.x ^ x.x) + __popc(a.y ^ x.y) + __popc(a.z ^ x.z) + __popc(a.w ^ x.w);
            else v += __popc(a.x & x.x) + __popc(a.y & x.y) + __popc(a.z & x.z) + __popc(a.w & x.w);
            reinterpret_cast<uint32_t &>(acc[q]) = v;
        }
    }
}

// acc_step2 argument order is (acc, q, x): L2/L1 compute q - x like the reference's a[i] - b[i] with a = query.
// (|q - x| and (q - x)^2 are symmetric, but keep the operand order anyway.)

size_t flat_smem_bytes(int elem, uint32_t row_stride, int k, FlatGeom *out) {
    FlatGeom g;
    g.chunk_bytes = row_stride < (uint32_t)FLAT_CB ? row_stride : (uint32_t)FLAT_CB;
    g.n_chunks = (int)((row_stride + g.chunk_bytes - 1) / g.chunk_bytes);
    g.rstride = g.chunk_bytes + (((g.chunk_bytes / 16) & 1u) ? 0u : 16u);
    g.q_chunk_bytes = elem == ELEM_F16 ? g.chunk_bytes * 2 : g.chunk_bytes;
    size_t rows = (size_t)FLAT_THREADS * g.rstride;
    size_t cand = (size_t)FLAT_QT * FLAT_THREADS * 8;
    g.rows_bytes = (uint32_t)(rows > cand ? rows : cand);
    g.rows_bytes = (g.rows_bytes + 127u) & ~127u;
    g.qs_bytes = (uint32_t)FLAT_QT * g.q_chunk_bytes;
    g.list_bytes = (uint32_t)(((size_t)FLAT_QT * k * 8 + 15) & ~(size_t)15);
    if (out) *out = g;
    // rows | query tile | lists | tau[QT] u64 | cnt[QT] | len[QT] | aa[QT] | mbar
    return (size_t)g.rows_bytes + g.qs_bytes + g.list_bytes + FLAT_QT * 8 + FLAT_QT * 4 * 3 + 16;
}

template <int ELEM, int METRIC>
__global__ void __launch_bounds__(FLAT_THREADS) flat_partial_kernel(FlatParams p, FlatGeom g) {
    extern __shared__ __align__(128) uint8_t smem[];
    typedef typename FlatAcc<ELEM>::type acc_t;
    uint8_t *rows = smem;
    uint64_t *cand = reinterpret_cast<uint64_t *>(smem);                       // aliases rows (used between tiles)
    uint4 *qs = reinterpret_cast<uint4 *>(smem + g.rows_bytes);
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem + g.rows_bytes + g.qs_bytes);
    uint64_t *tau = reinterpret_cast<uint64_t *>(smem + g.rows_bytes + g.qs_bytes + g.list_bytes);
    int *cnt = reinterpret_cast<int *>(tau + FLAT_QT);
    int *llen = cnt + FLAT_QT;
    uint32_t *aa = reinterpret_cast<uint32_t *>(llen + FLAT_QT);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(aa + FLAT_QT);

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int q0 = blockIdx.x * FLAT_QT;
    const int nq_here = min(FLAT_QT, p.nq - q0);
    const int64_t lo = p.row_begin + (int64_t)blockIdx.y * p.rows_per_split;
    const int64_t hi = p.row_end < lo + p.rows_per_split ? p.row_end : lo + p.rows_per_split;
    const int k = p.k;

    if (t < FLAT_QT) { tau[t] = FKEY_MAX; cnt[t] = 0; llen[t] = 0; aa[t] = 0; }
    if (t == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    __syncthreads();
    const uint64_t policy = l2_policy_evict_first();
    uint32_t phase = 0;
    const int nvec = (int)(g.chunk_bytes / 16);

    // stage dimension chunk c of the query tile, transposed to [vector][query]; halves are widened to f32
    auto stage_queries = [&](int c) {
        const uint32_t off = (uint32_t)c * g.chunk_bytes;
        const int qvec = (int)(g.q_chunk_bytes / 16);
        for (int i = t; i < qvec * FLAT_QT; i += FLAT_THREADS) {
            const int q = i % FLAT_QT, v = i / FLAT_QT;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (q < nq_here) {
                const uint8_t *src = p.queries + (size_t)(q0 + q) * p.q_stride;
                if (ELEM == ELEM_F16) {
                    // f32 vector v covers halves [4v, 4v + 4) of the chunk (rows of halves are 2-byte aligned)
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const uint32_t b = off + (uint32_t)(4 * v + e) * 2;
                        const uint16_t h = b + 1 < p.q_bytes ? __ldg(reinterpret_cast<const uint16_t *>(src + b)) : (uint16_t)0;
                        w[e] = __float_as_uint(__half2float(__ushort_as_half(h)));
                    }
                    val = make_uint4(w[0], w[1], w[2], w[3]);
                } else if (ELEM == ELEM_F32) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const uint32_t b = off + (uint32_t)v * 16 + (uint32_t)e * 4;
                        w[e] = b + 3 < p.q_bytes ? __ldg(reinterpret_cast<const uint32_t *>(src + b)) : 0u;
                    }
                    val = make_uint4(w[0], w[1], w[2], w[3]);
                } else {
                    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const uint32_t b = off + (uint32_t)v * 16 + e;
                        if (b < p.q_bytes) w[e >> 2] |= (uint32_t)__ldg(src + b) << (8 * (e & 3));
                    }
                    val = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            qs[(size_t)v * FLAT_QT + q] = val;
        }
    };

    if (g.n_chunks == 1) stage_queries(0);
    if (ELEM == ELEM_BIT && METRIC == METRIC_JACCARD) {
        // aa = popcount of each query (bitvec.rs:123), once per CTA
        if (t < nq_here) {
            const uint8_t *src = p.queries + (size_t)(q0 + t) * p.q_stride;
            uint32_t s = 0;
            for (uint32_t b = 0; b < p.q_bytes; b++) s += __popc((uint32_t)__ldg(src + b));
            aa[t] = s;
        }
    }
    __syncthreads();

    for (int64_t tile = lo; tile < hi; tile += FLAT_THREADS) {
        const int nrows = (int)((hi - tile) < (int64_t)FLAT_THREADS ? (hi - tile) : (int64_t)FLAT_THREADS);
        acc_t acc[FLAT_QT];
#pragma unroll
        for (int q = 0; q < FLAT_QT; q++) acc[q] = 0;
        uint32_t bb = 0;
        for (int c = 0; c < g.n_chunks; c++) {
            const uint32_t off = (uint32_t)c * g.chunk_bytes;
            const uint32_t cb = min(g.chunk_bytes, p.row_stride - off);
            if (g.n_chunks > 1) {
                __syncthreads();            // everyone is done with the previous chunk's rows and query tile
                stage_queries(c);
            }
            // generic-proxy accesses to the rows region (candidate buffers, LDS) are ordered before the TMA writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (t == 0) mbar_arrive_expect_tx(mbar, (uint32_t)nrows * cb);
            __syncthreads();
            if (t < nrows)
                tma_row_g2s(rows + (size_t)t * g.rstride, p.vectors + (size_t)(tile + t) * p.row_stride + off, cb, mbar, policy);
            mbar_wait(mbar, phase);
            phase ^= 1u;
            if (t < nrows) {
                const uint4 *r4 = reinterpret_cast<const uint4 *>(rows + (size_t)t * g.rstride);
                const int nv = (int)(cb / 16);
#pragma unroll 1
                for (int j = 0; j < nv; j++) flat_step<ELEM, METRIC>(acc, bb, r4[j], qs, j);
            }
        }
        __syncthreads();                    // rows region becomes the candidate buffers
        bool any = false;
        // a row whose element a vacuum has emptied or deleted is not in the table any more (its heap tuples are dead)
        if (t < nrows && !(p.flags != nullptr && __ldg(p.flags + (tile + t)) != 0)) {
            const uint32_t id = (uint32_t)(tile + t - p.row_begin) + (uint32_t)p.id_base;
#pragma unroll
            for (int q = 0; q < FLAT_QT; q++) {
                float d;
                if (ELEM == ELEM_BIT) {
                    const uint32_t v = reinterpret_cast<uint32_t &>(acc[q]);
                    if (METRIC == METRIC_HAMMING) d = (float)v;
                    else if (v == 0) d = 1.0f;                                               // bitvec.rs:127-128
                    else d = (float)(1.0 - ((double)v / (double)(aa[q] + bb - v)));          // bitvec.rs:130
                } else {
                    const float v = reinterpret_cast<float &>(acc[q]);
                    d = METRIC == METRIC_NEG_IP ? -v : v;
                }
                const uint64_t key = flat_key(d, id);
                if (q < nq_here && key < tau[q]) {
                    const int pos = atomicAdd(&cnt[q], 1);
                    cand[(size_t)q * FLAT_THREADS + pos] = key;
                    any = true;
                }
            }
        }
        if (__syncthreads_or(any)) {
            for (int q = warp; q < nq_here; q += FLAT_THREADS / 32) {
                const int n = cnt[q];
                if (n == 0) continue;
                int len = llen[q];
                uint64_t *L = lists + (size_t)q * k;
                for (int i = 0; i < n; i++) topk_insert(L, len, k, cand[(size_t)q * FLAT_THREADS + i], lane);
                if (lane == 0) { llen[q] = len; cnt[q] = 0; tau[q] = len == k ? L[k - 1] : FKEY_MAX; }
            }
        }
        // the loop-top fence + __syncthreads orders these reads/writes before the next tile's TMA
    }
    __syncthreads();
    for (int q = warp; q < nq_here; q += FLAT_THREADS / 32) {
        const int len = llen[q];
        const uint64_t *L = lists + (size_t)q * k;
        const size_t o = ((size_t)blockIdx.y * p.nq + (size_t)(q0 + q)) * k;
        for (int i = lane; i < k; i += 32) {
            if (i < len) { p.out_ids[o + i] = (int32_t)(uint32_t)L[i]; p.out_dists[o + i] = ord_f32((uint32_t)(L[i] >> 32)); }
            else { p.out_ids[o + i] = -1; p.out_dists[o + i] = __int_as_float(0x7f800000); }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K6: merge n_parts ascending (distance, id) lists of k_in entries per query into the k_out smallest.
// Parts are laid out [part][query][k_in] -- the layout an all-gather of per-shard results produces.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) topk_merge_kernel(const int32_t *__restrict__ ids, const float *__restrict__ dists,
                                                         int n_parts, int nq, int k_in, int k_out, int32_t *out_ids,
                                                         float *out_dists) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * 4 + warp;
    if (q >= nq) return;
    uint64_t *L = reinterpret_cast<uint64_t *>(smem) + (size_t)warp * k_out;
    int len = 0;
    for (int part = 0; part < n_parts; part++) {
        const size_t o = ((size_t)part * nq + q) * k_in;
        for (int base = 0; base < k_in; base += 32) {
            const int i = base + lane;
            uint64_t key = FKEY_MAX;
            if (i < k_in) {
                const int32_t id = ids[o + i];
                if (id >= 0) key = flat_key(dists[o + i], (uint32_t)id);
            }
            const uint64_t t = len == k_out ? L[k_out - 1] : FKEY_MAX;
            uint32_t am = __ballot_sync(FULL, key < t);
            if (!am) break;                     // lists ascend: nothing later in this part can qualify either
            while (am) {
                const int r = __ffs(am) - 1;
                am &= am - 1;
                topk_insert(L, len, k_out, __shfl_sync(FULL, key, r), lane);
            }
        }
    }
    __syncwarp();
    for (int i = lane; i < k_out; i += 32) {
        const size_t o = (size_t)q * k_out + i;
        if (i < len) { out_ids[o] = (int32_t)(uint32_t)L[i]; out_dists[o] = ord_f32((uint32_t)(L[i] >> 32)); }
        else { out_ids[o] = -1; out_dists[o] = __int_as_float(0x7f800000); }
    }
}

// The same merge with one pointer pair per part (parts may live on other GPUs: peer loads over NVLink).
__global__ void __launch_bounds__(128) topk_merge_parts_kernel(MergeParts parts, int nq, int k_in, int k_out, int32_t *out_ids,
                                                               float *out_dists) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * 4 + warp;
    if (q >= nq) return;
    uint64_t *L = reinterpret_cast<uint64_t *>(smem) + (size_t)warp * k_out;
    int len = 0;
    for (int part = 0; part < parts.n_parts; part++) {
        const int32_t *ids = parts.ids[part] + (size_t)q * k_in;
        const float *dists = parts.dists[part] + (size_t)q * k_in;
        for (int base = 0; base < k_in; base += 32) {
            const int i = base + lane;
            uint64_t key = FKEY_MAX;
            if (i < k_in) {
                const int32_t id = ids[i];
                if (id >= 0) key = flat_key(dists[i], (uint32_t)id);
            }
            const uint64_t t = len == k_out ? L[k_out - 1] : FKEY_MAX;
            uint32_t am = __ballot_sync(FULL, key < t);
            if (!am) break;                     // lists ascend: nothing later in this part can qualify either
            while (am) {
                const int r = __ffs(am) - 1;
                am &= am - 1;
                topk_insert(L, len, k_out, __shfl_sync(FULL, key, r), lane);
            }
        }
    }
    __syncwarp();
    for (int i = lane; i < k_out; i += 32) {
        const size_t o = (size_t)q * k_out + i;
        if (i < len) { out_ids[o] = (int32_t)(uint32_t)L[i]; out_dists[o] = ord_f32((uint32_t)(L[i] >> 32)); }
        else { out_ids[o] = -1; out_dists[o] = __int_as_float(0x7f800000); }
    }
}

// ---------------------------------------------------------------------------------------------------
// exact re-rank: per query, nc candidate row ids (-1 = none) -> exact distances, k smallest by (distance, id).
// One warp per query; rows are gathered with per-row TMA bulk copies and evaluated one lane per row, exactly
// like the scan kernel's neighbour evaluation.
// ---------------------------------------------------------------------------------------------------
template <int ELEM, int METRIC>
__global__ void __launch_bounds__(128) rerank_kernel(RerankParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per_warp = (size_t)p.R * p.stage_stride + p.row_stride + (((size_t)p.k * 8 + 15) & ~(size_t)15) + 256 + 16;
    uint8_t *base = smem + (size_t)warp * per_warp;
    WarpCtx c;
    c.stage = base;
    c.qsm = base + (size_t)p.R * p.stage_stride;
    uint64_t *L = reinterpret_cast<uint64_t *>(c.qsm + p.row_stride);
    c.nbr = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(L) + (((size_t)p.k * 8 + 15) & ~(size_t)15));
    c.mbar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(c.nbr) + 256);     // c.nbr holds up to 64 pending ids here
    c.stage_stride = p.stage_stride;
    c.R = p.R;
    c.chunk_bytes = c.chunk_stride = 0;
    c.nvec = (int)(p.row_stride / 16);
    c.lane = lane;
    c.phase = 0;
    c.phase2 = 0;
    c.stage2 = nullptr;
    c.policy = l2_policy_evict_first();
    GraphView g{};
    g.vectors = p.vectors;
    g.row_stride = p.row_stride;
    if (lane == 0) { mbar_init(c.mbar, 1); fence_mbar_init(); }
    __syncwarp();
    for (int q = blockIdx.x * 4 + warp; q < p.nq; q += gridDim.x * 4) {
        const uint8_t *qp = p.queries + (size_t)q * p.q_stride;
        for (uint32_t i = lane; i < p.row_stride; i += 32) c.qsm[i] = i < p.q_bytes ? __ldg(qp + i) : (uint8_t)0;
        __syncwarp();
        int len = 0;
        // candidate lists may be sparse (-1 padded segments): valid ids are collected across 32-id blocks and evaluated 32
        // at a time, so a gather round is always full
        int npend = 0;
        auto evaluate = [&](int cnt) {                         // the ids in c.nbr[0, cnt), cnt <= 32
            uint64_t key = FKEY_MAX;
            for (int s = 0; s < cnt; s += c.R) {
                const int nrow = min(c.R, cnt - s);
                gather_rows(c, g, s, nrow);
                if (lane >= s && lane < s + nrow) {
                    const float d = row_distance<ELEM, METRIC>(c.qsm, c.stage + (size_t)(lane - s) * c.stage_stride, c.nvec);
                    key = flat_key(d, (uint32_t)(c.nbr[lane] + p.id_base));
                }
                __syncwarp();
            }
            // only keys that beat the current k-th can enter (the list only tightens while they are inserted)
            uint32_t am = __ballot_sync(FULL, lane < cnt && (len < p.k || key < L[p.k - 1]));
            while (am) {
                const int r = __ffs(am) - 1;
                am &= am - 1;
                topk_insert(L, len, p.k, __shfl_sync(FULL, key, r), lane);
            }
        };
        for (int base_c = 0; base_c < p.nc; base_c += 32) {
            const int i = base_c + lane;
            const int32_t id = i < p.nc ? p.cand[(size_t)q * p.nc + i] : -1;
            bool valid = id >= 0 && (int64_t)id - p.id_base < p.n_rows && id >= p.id_base;
            if (valid && p.flags != nullptr && __ldg(p.flags + (id - p.id_base)) != 0) valid = false;
            const uint32_t vm = __ballot_sync(FULL, valid);
            if (!vm) continue;
            const int slot = npend + __popc(vm & ((1u << lane) - 1u));
            if (valid) c.nbr[slot] = id - p.id_base;
            npend += __popc(vm);
            __syncwarp();
            if (npend >= 32) {
                evaluate(32);
                const int32_t carry = lane < npend - 32 ? c.nbr[32 + lane] : 0;
                __syncwarp();
                if (lane < npend - 32) c.nbr[lane] = carry;
                npend -= 32;
                __syncwarp();
            }
        }
        if (npend > 0) evaluate(npend);
        for (int i = lane; i < p.k; i += 32) {
            const size_t o = (size_t)q * p.k + i;
            if (i < len) { p.out_ids[o] = (int32_t)(uint32_t)L[i]; p.out_dists[o] = ord_f32((uint32_t)(L[i] >> 32)); }
            else { p.out_ids[o] = -1; p.out_dists[o] = __int_as_float(0x7f800000); }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
typedef void (*flat_kernel_t)(FlatParams, FlatGeom);
typedef void (*rerank_kernel_t)(RerankParams);

static flat_kernel_t pick_flat(int elem, int metric) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return flat_partial_kernel<ELEM_F32, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return flat_partial_kernel<ELEM_F32, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return flat_partial_kernel<ELEM_F32, METRIC_L1>;
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return flat_partial_kernel<ELEM_F16, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return flat_partial_kernel<ELEM_F16, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return flat_partial_kernel<ELEM_F16, METRIC_L1>;
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return flat_partial_kernel<ELEM_BIT, METRIC_HAMMING>;
        if (metric == METRIC_JACCARD) return flat_partial_kernel<ELEM_BIT, METRIC_JACCARD>;
    }
    return nullptr;
}

static rerank_kernel_t pick_rerank(int elem, int metric) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return rerank_kernel<ELEM_F32, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return rerank_kernel<ELEM_F32, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return rerank_kernel<ELEM_F32, METRIC_L1>;
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return rerank_kernel<ELEM_F16, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return rerank_kernel<ELEM_F16, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return rerank_kernel<ELEM_F16, METRIC_L1>;
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return rerank_kernel<ELEM_BIT, METRIC_HAMMING>;
        if (metric == METRIC_JACCARD) return rerank_kernel<ELEM_BIT, METRIC_JACCARD>;
    }
    return nullptr;
}

cudaError_t flat_kernel_occupancy(int elem, int metric, size_t smem, int *blocks_per_sm) {
    flat_kernel_t kfn = pick_flat(elem, metric);
    if (!kfn) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kfn, FLAT_THREADS, smem);
}

cudaError_t launch_flat_partial(int elem, int metric, const FlatParams &p, const FlatGeom &g, int n_qtiles, int n_splits,
                                size_t smem, cudaStream_t stream) {
    flat_kernel_t kfn = pick_flat(elem, metric);
    if (!kfn) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kfn<<<dim3((unsigned)n_qtiles, (unsigned)n_splits), FLAT_THREADS, smem, stream>>>(p, g);
    return cudaGetLastError();
}

cudaError_t launch_topk_merge(const int32_t *ids, const float *dists, int n_parts, int nq, int k_in, int k_out,
                              int32_t *out_ids, float *out_dists, cudaStream_t stream) {
    const size_t smem = (size_t)4 * k_out * 8;
    cudaError_t e = cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    topk_merge_kernel<<<(nq + 3) / 4, 128, smem, stream>>>(ids, dists, n_parts, nq, k_in, k_out, out_ids, out_dists);
    return cudaGetLastError();
}

cudaError_t launch_topk_merge_parts(const MergeParts &parts, int nq, int k_in, int k_out, int32_t *out_ids, float *out_dists,
                                    cudaStream_t stream) {
    const size_t smem = (size_t)4 * k_out * 8;
    cudaError_t e = cudaFuncSetAttribute(topk_merge_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    topk_merge_parts_kernel<<<(nq + 3) / 4, 128, smem, stream>>>(parts, nq, k_in, k_out, out_ids, out_dists);
    return cudaGetLastError();
}

size_t rerank_smem_bytes(uint32_t row_stride, uint32_t stage_stride, int R, int k) {
    return 4 * ((size_t)R * stage_stride + row_stride + (((size_t)k * 8 + 15) & ~(size_t)15) + 256 + 16);
}

cudaError_t launch_rerank(int elem, int metric, const RerankParams &p, int grid, size_t smem, cudaStream_t stream) {
    rerank_kernel_t kfn = pick_rerank(elem, metric);
    if (!kfn) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kfn<<<grid, 128, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace pgvrx
