// seam.cu -- the narrow build seam of SURVEY 8(b)(4)/(5): a caller that keeps select_neighbors /
// update_neighbor_connections / the insert loop on ITS side (the pgrx build_callback, src/index/build.rs:400-535) and
// only wants the GPU where the time goes -- the search_layer calls of find_element_neighbors
// (src/graph/mod.rs:386,407) -- binds two entry points:
//   pgvrx_search_layer_batch   search_layer(query, entry_points, ef, layer) for a batch of queries against the resident
//                              graph: W sorted by (distance, id), plus the ids of the elements each search expanded
//                              (its read set: what the caller needs to decide whether a speculative result survives the
//                              commits made since the snapshot);
//   pgvrx_commit_insert        the caller's decision for ONE new element -- its level, row, neighbour lists and the
//                              neighbour lists it rewrote -- applied to the resident graph, so the next searches see it.
// pgvrx_index_build / pgvrx_index_insert are the same machinery with the heuristic and the commit inside the library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "hnsw_device.cuh"
#include "index_internal.h"

namespace pgvrx {

struct LayerSearchParams {
    const uint8_t *queries;    // nq rows, q_stride apart, or NULL
    uint32_t q_stride, q_bytes;
    const int32_t *query_ids;  // nq element ids whose stored rows are the queries, or NULL
    int nq, ef, tail_cap, layer;
    const int32_t *ep_ids;     // nq x ep_pitch
    const float *ep_dists;
    const int32_t *ep_counts;  // nq
    int ep_pitch;
    int32_t *out_ids;          // nq x ef
    float *out_dists;
    int32_t *out_counts;
    int32_t *rs_log;           // nq x rs_cap x rs_ent
    int rs_cap, rs_ent;
    int32_t *rs_counts;
    uint32_t *out_status;
    uint32_t *vis_base;
    uint32_t vis_entries;
    int *work_counter;
    uint32_t stage_stride;
    int R;
    uint32_t stage_bytes, chunk_bytes, chunk_stride;
};

__host__ __device__ inline uint32_t round16s(uint32_t x) { return (x + 15u) & ~15u; }

// one warp per query, result list in shared memory (any ef)
template <int ELEM, int METRIC>
__global__ void __launch_bounds__(32) hnsw_layer_search_kernel(GraphView g, LayerSearchParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int cap = p.ef + p.tail_cap;
    WarpCtx c;
    c.stage = smem;
    c.qsm = c.stage + p.stage_bytes;
    uint8_t *s_mem = c.qsm + g.row_stride;
    const size_t s_bytes = round16s((uint32_t)((cap + 1) * 8));
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
    c.stage2 = nullptr;
    c.policy = l2_policy_evict_first();
    c.spec = true;
    c.hybrid = false;
    c.vis = p.vis_base + (size_t)blockIdx.x * p.vis_entries;
    if (lane == 0) { mbar_init(c.mbar, 1); fence_mbar_init(); }
    __syncwarp();
    while (true) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.work_counter, 1);
        w = __shfl_sync(FULL, w, 0);
        if (w >= p.nq) break;
        c.n_dist = c.n_exp0 = c.n_expu = c.n_vis = 0;
        c.status = 0;
        c.rs_log = p.rs_log + (size_t)w * p.rs_cap * p.rs_ent;
        c.rs_ent = p.rs_ent;
        c.rs_cap = p.rs_cap;
        c.rs_n = 0;
        const uint8_t *q = p.query_ids ? g.vectors + (size_t)p.query_ids[w] * g.row_stride : p.queries + (size_t)w * p.q_stride;
        const uint32_t qb = p.query_ids ? g.row_stride : p.q_bytes;
        for (uint32_t i = lane; i < g.row_stride; i += 32) c.qsm[i] = i < qb ? __ldg(q + i) : (uint8_t)0;
        // entry points: visited, in C and in W, nothing evicted while seeding (graph/mod.rs:171-184)
        vis_clear(c.vis, p.vis_entries, lane);
        c.ef = p.ef;
        c.len = 0;
        c.f_ord = 0;
        const int n_ep = p.ep_counts[w];
        for (int i = 0; i < n_ep; i++) {
            const int32_t id = p.ep_ids[(size_t)w * p.ep_pitch + i];
            const float d = p.ep_dists[(size_t)w * p.ep_pitch + i];
            if (lane == 0) vis_insert(c.vis, p.vis_entries - 1, (uint32_t)id);
            __syncwarp();
            S.insert(c, make_key(d, (uint32_t)id));
        }
        c.n_vis = (uint32_t)n_ep;
        __syncwarp();
        if (n_ep > 0) search_layer<ELEM, METRIC>(c, S, g, p.layer, p.vis_entries - 1, p.vis_entries / 4 * 3 - 32);
        const int out_n = c.status == 0 ? min(c.len, c.ef) : 0;
        S.store(c, p.out_ids + (size_t)w * p.ef, p.out_dists + (size_t)w * p.ef, out_n);
        for (int i = out_n + lane; i < p.ef; i += 32) {
            p.out_ids[(size_t)w * p.ef + i] = -1;
            p.out_dists[(size_t)w * p.ef + i] = __int_as_float(0x7f800000);
        }
        if (lane == 0) {
            p.out_counts[w] = c.status == 0 ? out_n : -1;
            p.out_status[w] = c.status;
            p.rs_counts[w] = c.rs_n;
        }
        __syncwarp();
    }
}

typedef void (*layer_kernel_t)(GraphView, LayerSearchParams);
static layer_kernel_t pick_layer_kernel(int elem, int metric) {
    if (elem == ELEM_F32) {
        if (metric == METRIC_L2SQ) return hnsw_layer_search_kernel<ELEM_F32, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return hnsw_layer_search_kernel<ELEM_F32, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return hnsw_layer_search_kernel<ELEM_F32, METRIC_L1>;
    } else if (elem == ELEM_F16) {
        if (metric == METRIC_L2SQ) return hnsw_layer_search_kernel<ELEM_F16, METRIC_L2SQ>;
        if (metric == METRIC_NEG_IP) return hnsw_layer_search_kernel<ELEM_F16, METRIC_NEG_IP>;
        if (metric == METRIC_L1) return hnsw_layer_search_kernel<ELEM_F16, METRIC_L1>;
    } else if (elem == ELEM_BIT) {
        if (metric == METRIC_HAMMING) return hnsw_layer_search_kernel<ELEM_BIT, METRIC_HAMMING>;
        if (metric == METRIC_JACCARD) return hnsw_layer_search_kernel<ELEM_BIT, METRIC_JACCARD>;
    }
    return nullptr;
}

}  // namespace pgvrx

using namespace pgvrx;

extern "C" {

int pgvrx_search_layer_batch(pgvrx_index *ix, const void *queries, const int32_t *query_ids, int64_t nq, const int32_t *ep_ids,
                             const float *ep_dists, const int32_t *ep_counts, int32_t ep_pitch, int32_t ef, int32_t layer,
                             int32_t *out_ids, float *out_dists, int32_t *out_counts, int32_t *out_expanded, int32_t exp_cap,
                             int32_t *out_n_expanded) try {
    if (!ix) return PGVRX_ERR_INVALID;
    if (!ix->loaded) return pgvrx_fail(ix, PGVRX_ERR_STATE, "no graph loaded");
    if (nq < 0 || nq > (int64_t)1 << 24) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad nq");
    if (ef < 1 || ef > PGVRX_MAX_EF_CONSTRUCTION) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "ef out of range (1 .. 1000)");
    if (layer < 0 || layer > 255) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad layer");
    if ((queries == nullptr) == (query_ids == nullptr)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "give either query rows or query element ids");
    if (nq > 0 && (!ep_ids || !ep_dists || !ep_counts || ep_pitch < 1 || !out_ids || !out_dists || !out_counts))
        return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null buffer");
    if (exp_cap < 0 || (exp_cap > 0 && (!out_expanded || !out_n_expanded))) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad expanded-set buffer");
    if (nq == 0) return PGVRX_OK;
    // entry points are trusted by the kernel: validate them here (ids in range and on the layer, at most ef of them)
    std::vector<uint8_t> h_level;
    for (int64_t i = 0; i < nq; i++) {
        if (ep_counts[i] < 0 || ep_counts[i] > ep_pitch || ep_counts[i] > ef)
            return pgvrx_fail(ix, PGVRX_ERR_INVALID, "entry point count must be within 0 .. min(ep_pitch, ef)");
        for (int t = 0; t < ep_counts[i]; t++) {
            const int32_t id = ep_ids[(size_t)i * ep_pitch + t];
            if (id < 0 || id >= ix->n) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "entry point id out of range");
        }
        if (query_ids && (query_ids[i] < 0 || query_ids[i] >= ix->n)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "query element id out of range");
    }
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    cudaStream_t s = ix->stream;
    const int m = ix->desc.m, lm0 = 2 * m;
    if (layer > 0) {
        // an entry point below `layer` has no list there (the reference skips it, graph/mod.rs:198; here it is the
        // caller's error: entry points come from the layer above)
        h_level.resize((size_t)ix->n);
        CU_TRY(ix, cudaMemcpy(h_level.data(), ix->level.p, (size_t)ix->n, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < nq; i++)
            for (int t = 0; t < ep_counts[i]; t++)
                if (h_level[(size_t)ep_ids[(size_t)i * ep_pitch + t]] < layer) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "entry point is not on the searched layer");
    }
    ScanConfig cfg;
    const uint32_t vis_entries = pgvrx_pow2_at_least(std::min<uint64_t>(std::max<uint64_t>(1024, (uint64_t)ef * 2 * m * 8), std::max<uint64_t>(1024, (uint64_t)ix->n * 2)));
    const int tail_cap = std::max(64, std::min(1023, ef));         // retained boundary ties: at most ef - 1 (see build_host.cu)
    int rc = pgvrx_make_config(ix, ef, tail_cap, vis_entries, nq, 0, true, &cfg);
    if (rc) return rc;
    layer_kernel_t kfn = pick_layer_kernel(ix->desc.elem_type, ix->desc.metric);
    if (!kfn) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "operator class does not match the indexed type");
    SearchLaunch L = cfg.L;
    L.list_j = 0;
    const size_t smem = search_warp_smem(L);
    if (smem > 227 * 1024) return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "search configuration exceeds shared memory");
    CU_TRY(ix, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CU_TRY(ix, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, 32, smem));
    if (occ < 1) return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, "search kernel does not fit on an SM");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nq, (int64_t)ix->sm_count * occ));
    const int rs_cap = std::max(256, 6 * ef + 64), rs_ent = 3 + (lm0 + 31) / 32;
    DevBuf d_q, d_qid, d_ep, d_epd, d_epc, d_ids, d_dists, d_cnt, d_rs, d_rsc, d_st, d_vis, d_work;
    struct Rel { DevBuf *b[13]; ~Rel() { for (DevBuf *x : b) x->release(); } } rel{{&d_q, &d_qid, &d_ep, &d_epd, &d_epc, &d_ids, &d_dists, &d_cnt, &d_rs, &d_rsc, &d_st, &d_vis, &d_work}};
    if (queries) CU_TRY(ix, d_q.ensure((size_t)nq * ix->row_bytes)); else CU_TRY(ix, d_qid.ensure((size_t)nq * 4));
    CU_TRY(ix, d_ep.ensure((size_t)nq * ep_pitch * 4));
    CU_TRY(ix, d_epd.ensure((size_t)nq * ep_pitch * 4));
    CU_TRY(ix, d_epc.ensure((size_t)nq * 4));
    CU_TRY(ix, d_ids.ensure((size_t)nq * ef * 4));
    CU_TRY(ix, d_dists.ensure((size_t)nq * ef * 4));
    CU_TRY(ix, d_cnt.ensure((size_t)nq * 4));
    CU_TRY(ix, d_rs.ensure((size_t)nq * rs_cap * rs_ent * 4));
    CU_TRY(ix, d_rsc.ensure((size_t)nq * 4));
    CU_TRY(ix, d_st.ensure((size_t)nq * 4));
    CU_TRY(ix, d_vis.ensure((size_t)grid * vis_entries * 4));
    CU_TRY(ix, d_work.ensure(64));
    if (queries) CU_TRY(ix, cudaMemcpyAsync(d_q.p, queries, (size_t)nq * ix->row_bytes, cudaMemcpyHostToDevice, s));
    else CU_TRY(ix, cudaMemcpyAsync(d_qid.p, query_ids, (size_t)nq * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(ix, cudaMemcpyAsync(d_ep.p, ep_ids, (size_t)nq * ep_pitch * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(ix, cudaMemcpyAsync(d_epd.p, ep_dists, (size_t)nq * ep_pitch * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(ix, cudaMemcpyAsync(d_epc.p, ep_counts, (size_t)nq * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(ix, cudaMemsetAsync(d_work.p, 0, 64, s));
    LayerSearchParams p;
    p.queries = queries ? (const uint8_t *)d_q.p : nullptr;
    p.q_stride = (uint32_t)ix->row_bytes;
    p.q_bytes = (uint32_t)ix->row_bytes;
    p.query_ids = queries ? nullptr : (const int32_t *)d_qid.p;
    p.nq = (int)nq;
    p.ef = ef;
    p.tail_cap = tail_cap;
    p.layer = layer;
    p.ep_ids = (const int32_t *)d_ep.p;
    p.ep_dists = (const float *)d_epd.p;
    p.ep_counts = (const int32_t *)d_epc.p;
    p.ep_pitch = ep_pitch;
    p.out_ids = (int32_t *)d_ids.p;
    p.out_dists = (float *)d_dists.p;
    p.out_counts = (int32_t *)d_cnt.p;
    p.rs_log = (int32_t *)d_rs.p;
    p.rs_cap = rs_cap;
    p.rs_ent = rs_ent;
    p.rs_counts = (int32_t *)d_rsc.p;
    p.out_status = (uint32_t *)d_st.p;
    p.vis_base = (uint32_t *)d_vis.p;
    p.vis_entries = vis_entries;
    p.work_counter = (int *)d_work.p;
    p.stage_stride = L.stage_stride;
    p.R = L.R;
    p.stage_bytes = L.stage_bytes;
    p.chunk_bytes = L.chunk_bytes;
    p.chunk_stride = L.chunk_stride;
    kfn<<<grid, 32, smem, s>>>(pgvrx_view_of(ix), p);
    CU_TRY(ix, cudaGetLastError());
    CU_TRY(ix, cudaMemcpyAsync(out_ids, d_ids.p, (size_t)nq * ef * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ix, cudaMemcpyAsync(out_dists, d_dists.p, (size_t)nq * ef * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ix, cudaMemcpyAsync(out_counts, d_cnt.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    std::vector<int32_t> h_rs, h_rsc((size_t)nq);
    std::vector<uint32_t> h_st((size_t)nq);
    if (exp_cap > 0) {
        h_rs.resize((size_t)nq * rs_cap * rs_ent);
        CU_TRY(ix, cudaMemcpyAsync(h_rs.data(), d_rs.p, h_rs.size() * 4, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(ix, cudaMemcpyAsync(h_rsc.data(), d_rsc.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ix, cudaMemcpyAsync(h_st.data(), d_st.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ix, cudaStreamSynchronize(s));
    for (int64_t i = 0; i < nq; i++) {
        if (h_st[(size_t)i] != 0) {
            char b[160];
            snprintf(b, sizeof b, "search %lld exhausted the search scratch (status %u)", (long long)i, h_st[(size_t)i]);
            return pgvrx_fail(ix, PGVRX_ERR_UNSUPPORTED, b);
        }
        if (exp_cap > 0) {
            // expansion order; when a search expanded more elements than fit, the count says so and the list is its prefix
            out_n_expanded[i] = h_rsc[(size_t)i];
            const int keep = std::min(std::min(h_rsc[(size_t)i], rs_cap), exp_cap);
            for (int t = 0; t < keep; t++) out_expanded[(size_t)i * exp_cap + t] = h_rs[((size_t)i * rs_cap + t) * rs_ent];
            for (int t = keep; t < exp_cap; t++) out_expanded[(size_t)i * exp_cap + t] = -1;
        }
    }
    return PGVRX_OK;
} catch (const std::bad_alloc &) {
    return ix ? pgvrx_fail(ix, PGVRX_ERR_NOMEM, "out of host memory") : PGVRX_ERR_NOMEM;
}

int pgvrx_commit_insert(pgvrx_index *ix, const void *row, int32_t level, const int32_t *lists, int32_t n_modified,
                        const int32_t *mod_nodes, const int32_t *mod_layers, const int32_t *mod_lists, int64_t *new_id_out) try {
    if (!ix) return PGVRX_ERR_INVALID;
    const int m = ix->desc.m, l0 = 2 * m;
    if (!row || level < 0 || level > pgvrx_max_level(m)) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad row or level");
    if (!lists) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "null neighbour lists");
    if (n_modified < 0 || (n_modified > 0 && (!mod_nodes || !mod_layers || !mod_lists))) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "bad modified lists");
    if (ix->has_flags) return pgvrx_fail(ix, PGVRX_ERR_STATE, "element flags are set (vacuumed index)");
    const bool empty = !ix->loaded || ix->n == 0;
    const int64_t e = empty ? 0 : ix->n;
    if (e + 1 > ix->desc.capacity) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "commit exceeds the handle's capacity");
    // the new element's lists: layer 0 first (2m ids), then layers 1 .. level (m ids each), -1 after the last neighbour
    const size_t list_ints = (size_t)l0 + (size_t)level * m;
    for (size_t t = 0; t < list_ints; t++)
        if (lists[t] < -1 || lists[t] >= e) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "neighbour id out of range in the new element's lists");
    size_t mod_ints = 0;
    for (int t = 0; t < n_modified; t++) {
        if (mod_nodes[t] < 0 || mod_nodes[t] >= e || mod_layers[t] < 0) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "modified list: node or layer out of range");
        mod_ints += mod_layers[t] == 0 ? (size_t)l0 : (size_t)m;
    }
    for (size_t t = 0; t < mod_ints; t++)
        if (mod_lists[t] < -1 || mod_lists[t] > e) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "neighbour id out of range in a modified list");
    CU_TRY(ix, cudaSetDevice(ix->desc.device));
    cudaStream_t s = ix->stream;
    const int64_t cap = ix->desc.capacity;
    const int64_t n_upper_old = empty ? 0 : ix->n_upper;
    if (empty) { ix->n = 0; ix->n_upper = 0; ix->entry = -1; ix->entry_level = -1; }
    CU_TRY(ix, ix->vectors.ensure_keep(std::max<size_t>(16, (size_t)cap * ix->row_stride), (size_t)e * ix->row_stride));
    CU_TRY(ix, ix->adj0.ensure_keep(std::max<size_t>(16, (size_t)cap * l0 * 4), (size_t)e * l0 * 4));
    CU_TRY(ix, ix->upper_base.ensure_keep(std::max<size_t>(16, (size_t)cap * 4), (size_t)e * 4));
    CU_TRY(ix, ix->level.ensure_keep(std::max<size_t>(16, (size_t)cap), (size_t)e));
    // upper rows grow geometrically (their total is not known in advance)
    const size_t up_need = (size_t)(n_upper_old + level) * m * 4;
    if (up_need > ix->adj_up.cap) CU_TRY(ix, ix->adj_up.ensure_keep(std::max<size_t>(up_need * 2, (size_t)1 << 16), (size_t)n_upper_old * m * 4));
    // modified levels must exist on their nodes: checked against the device's level array (small D2H, once per call)
    if (n_modified > 0) {
        std::vector<uint8_t> lv((size_t)e);
        CU_TRY(ix, cudaMemcpy(lv.data(), ix->level.p, (size_t)e, cudaMemcpyDeviceToHost));
        for (int t = 0; t < n_modified; t++)
            if (mod_layers[t] > lv[(size_t)mod_nodes[t]]) return pgvrx_fail(ix, PGVRX_ERR_INVALID, "modified list: layer above the node's level");
    }
    ix->generation++;
    ix->tc_ready = false;
    if (ix->build) { pgvrx_host::destroy_build_state(ix->build); ix->build = nullptr; }   // the host mirror no longer describes the graph
    // the element: row (zero padded to the row stride), level, first upper row, lists
    std::vector<uint8_t> padded(ix->row_stride, 0);
    memcpy(padded.data(), row, ix->row_bytes);
    const uint8_t lv8 = (uint8_t)level;
    const int32_t ub = level > 0 ? (int32_t)n_upper_old : -1;
    CU_TRY(ix, cudaMemcpyAsync((uint8_t *)ix->vectors.p + (size_t)e * ix->row_stride, padded.data(), ix->row_stride, cudaMemcpyHostToDevice, s));
    CU_TRY(ix, cudaMemcpyAsync((uint8_t *)ix->level.p + e, &lv8, 1, cudaMemcpyHostToDevice, s));
    CU_TRY(ix, cudaMemcpyAsync((int32_t *)ix->upper_base.p + e, &ub, 4, cudaMemcpyHostToDevice, s));
    CU_TRY(ix, cudaMemcpyAsync((int32_t *)ix->adj0.p + (size_t)e * l0, lists, (size_t)l0 * 4, cudaMemcpyHostToDevice, s));
    if (level > 0)
        CU_TRY(ix, cudaMemcpyAsync((int32_t *)ix->adj_up.p + (size_t)n_upper_old * m, lists + l0, (size_t)level * m * 4, cudaMemcpyHostToDevice, s));
    // the neighbours' rewritten lists (update_neighbor_connections, graph/mod.rs:442-489)
    if (n_modified > 0) {
        std::vector<int32_t> h_ub((size_t)e);
        bool need_ub = false;
        for (int t = 0; t < n_modified; t++) need_ub = need_ub || mod_layers[t] > 0;
        if (need_ub) CU_TRY(ix, cudaMemcpy(h_ub.data(), ix->upper_base.p, (size_t)e * 4, cudaMemcpyDeviceToHost));
        const int32_t *src = mod_lists;
        for (int t = 0; t < n_modified; t++) {
            const int lc = mod_layers[t];
            const size_t len = lc == 0 ? (size_t)l0 : (size_t)m;
            int32_t *dst = lc == 0 ? (int32_t *)ix->adj0.p + (size_t)mod_nodes[t] * l0
                                   : (int32_t *)ix->adj_up.p + ((size_t)h_ub[(size_t)mod_nodes[t]] + (size_t)(lc - 1)) * m;
            CU_TRY(ix, cudaMemcpyAsync(dst, src, len * 4, cudaMemcpyHostToDevice, s));
            src += len;
        }
    }
    CU_TRY(ix, cudaStreamSynchronize(s));                  // the caller's buffers may be reused; later searches see the commit
    ix->n = e + 1;
    ix->n_upper = n_upper_old + level;
    if (ix->entry < 0 || level > ix->entry_level) {          // build.rs:523-529: first element, or strictly greater level
        ix->entry = e;
        ix->entry_level = level;
    }
    ix->loaded = true;
    ix->rows_loaded = true;
    if (new_id_out) *new_id_out = e;
    return PGVRX_OK;
} catch (const std::bad_alloc &) {
    return ix ? pgvrx_fail(ix, PGVRX_ERR_NOMEM, "out of host memory") : PGVRX_ERR_NOMEM;
}

}  // extern "C"
