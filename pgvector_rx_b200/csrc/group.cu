// group.cu -- several GPUs of one box behind one handle (SURVEY 8e, 8b(1) "device list").
//
// One process, one host thread per device while a call runs (the reference's one-backend-one-thread model applied per
// GPU, build.rs:380-385), every device with its own pgvrx_index.  Two partitionings, as the paths shard:
//   * scan: the graph and the vector table are REPLICATED; one caller batch is split by queries, every device writes
//     its slice of the caller's result arrays (no exchange step: strong scaling of one batch through one call);
//   * exact flat kNN: the base rows are SHARDED contiguously; every device computes the top-k of its shard for all
//     queries with global ids, and the lead device merges the per-shard lists with ONE kernel that reads the peers'
//     lists directly over NVLink (peer access) -- the gather and the k-way merge are the same pass, nothing is staged.
//     Where peer access is not available the lists are copied to the lead device first (still inside this call).
// The multi-process form of the same flat path (one rank per GPU, NCCL all-gather + pgvrx_topk_merge_device) lives in
// pgvector_rx_b200/sharded.py for launchers that run one process per GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "index_internal.h"

using namespace pgvrx;

struct pgvrx_group {
    pgvrx_index_desc desc{};
    std::vector<pgvrx_index *> ix;
    std::vector<int> dev;
    std::vector<int64_t> lo, hi;             // row shard of every member (flat path)
    std::vector<char> peer;                  // lead device can read member i's memory directly
    int64_t n_rows = 0;
    bool rows_sharded = false, graph_replicated = false;
    struct Member { DevBuf q, ids, dists; };
    std::vector<Member> mem;
    DevBuf lead_ids, lead_dists, stage_ids, stage_dists;
    std::vector<cudaEvent_t> done;           // per member: its local top-k is complete
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
};

namespace {

thread_local std::string g_group_create_error;

int gfail(pgvrx_group *g, int code, const std::string &msg) {
    if (g) g->err = msg; else g_group_create_error = msg;
    return code;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// run fn(i) for every member on its own host thread; returns the first non-zero status
template <class F> int for_members(pgvrx_group *g, F fn) {
    const int n = (int)g->ix.size();
    std::vector<int> rc((size_t)n, 0);
    if (n == 1) { rc[0] = fn(0); }
    else {
        std::vector<std::thread> ts;
        int started = 0;
        try {
            ts.reserve((size_t)n);
            for (; started < n; started++) ts.emplace_back([&, started]() { rc[(size_t)started] = fn(started); });
        } catch (...) {
        }
        for (int i = started; i < n; i++) rc[(size_t)i] = fn(i);      // a thread that could not be created: run it here
        for (auto &t : ts) t.join();
    }
    for (int i = 0; i < n; i++)
        if (rc[(size_t)i]) {
            g->err = std::string("device ") + std::to_string(g->dev[(size_t)i]) + ": " + pgvrx_last_error(g->ix[(size_t)i]);
            return rc[(size_t)i];
        }
    return PGVRX_OK;
}

#define G_TRY(g, expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            std::string _m = std::string(cudaGetErrorName(_e)) + ": " + cudaGetErrorString(_e);         \
            cudaGetLastError();                                                                         \
            return gfail(g, _e == cudaErrorMemoryAllocation ? PGVRX_ERR_NOMEM : PGVRX_ERR_CUDA, _m);    \
        }                                                                                               \
    } while (0)

}  // namespace

extern "C" {

const char *pgvrx_group_last_error(const pgvrx_group *g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }

int pgvrx_group_create(const pgvrx_index_desc *desc, const int32_t *devices, int32_t n_devices, pgvrx_group **out) try {
    if (!desc || !devices || !out) return gfail(nullptr, PGVRX_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) return gfail(nullptr, PGVRX_ERR_INVALID, "a group holds 1 .. 64 devices");
    pgvrx_group *g = new (std::nothrow) pgvrx_group();
    if (!g) return gfail(nullptr, PGVRX_ERR_NOMEM, "out of host memory");
    g->desc = *desc;
    for (int i = 0; i < n_devices; i++) {
        pgvrx_index_desc d = *desc;
        d.device = devices[i];
        pgvrx_index *ix = nullptr;
        int rc = pgvrx_index_create(&d, &ix);
        if (rc) {
            std::string msg = pgvrx_last_error(nullptr);
            pgvrx_group_destroy(g);
            return gfail(nullptr, rc, msg);
        }
        g->ix.push_back(ix);
        g->dev.push_back(devices[i]);
    }
    g->mem.resize((size_t)n_devices);
    g->lo.assign((size_t)n_devices, 0);
    g->hi.assign((size_t)n_devices, 0);
    g->peer.assign((size_t)n_devices, 0);
    g->done.assign((size_t)n_devices, nullptr);
    // the lead device (member 0) reads the other members' result lists in place when the hardware allows it
    cudaError_t e = cudaSetDevice(g->dev[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&g->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&g->ev1);
    for (int i = 0; i < n_devices && e == cudaSuccess; i++) {
        if (g->dev[(size_t)i] == g->dev[0]) { g->peer[(size_t)i] = 1; continue; }
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, g->dev[0], g->dev[(size_t)i]) == cudaSuccess && can) {
            cudaError_t pe = cudaDeviceEnablePeerAccess(g->dev[(size_t)i], 0);
            if (pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled) g->peer[(size_t)i] = 1;
        }
        cudaGetLastError();
    }
    for (int i = 0; i < n_devices && e == cudaSuccess; i++) {
        e = cudaSetDevice(g->dev[(size_t)i]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->done[(size_t)i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        std::string msg = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e);
        cudaGetLastError();
        pgvrx_group_destroy(g);
        return gfail(nullptr, PGVRX_ERR_CUDA, msg);
    }
    *out = g;
    return PGVRX_OK;
} catch (const std::bad_alloc &) {
    return gfail(nullptr, PGVRX_ERR_NOMEM, "out of host memory");
}

void pgvrx_group_destroy(pgvrx_group *g) {
    if (!g) return;
    for (size_t i = 0; i < g->ix.size(); i++) {
        cudaSetDevice(g->dev[i]);
        if (i < g->mem.size()) { g->mem[i].q.release(); g->mem[i].ids.release(); g->mem[i].dists.release(); }
        if (i < g->done.size() && g->done[i]) cudaEventDestroy(g->done[i]);
        pgvrx_index_destroy(g->ix[i]);
    }
    if (!g->dev.empty()) {
        cudaSetDevice(g->dev[0]);
        g->lead_ids.release(); g->lead_dists.release(); g->stage_ids.release(); g->stage_dists.release();
        if (g->ev0) cudaEventDestroy(g->ev0);
        if (g->ev1) cudaEventDestroy(g->ev1);
    }
    delete g;
}

int32_t pgvrx_group_size(const pgvrx_group *g) { return g ? (int32_t)g->ix.size() : 0; }

pgvrx_index *pgvrx_group_member(pgvrx_group *g, int32_t i) {
    return (g && i >= 0 && i < (int32_t)g->ix.size()) ? g->ix[(size_t)i] : nullptr;
}

int pgvrx_group_peer_access(const pgvrx_group *g, int32_t i) {
    return (g && i >= 0 && i < (int32_t)g->peer.size()) ? g->peer[(size_t)i] : 0;
}

// ---- replicated graph: query-sharded scan ------------------------------------------------------------------------
int pgvrx_group_load_graph(pgvrx_group *g, const void *rows, int64_t n, const uint8_t *level, const int32_t *adj0,
                           const int32_t *upper_base, const int32_t *adj_up, int64_t n_upper_rows, int64_t entry) try {
    if (!g) return PGVRX_ERR_INVALID;
    int rc = for_members(g, [&](int i) { return pgvrx_index_load_graph(g->ix[(size_t)i], rows, n, level, adj0, upper_base, adj_up, n_upper_rows, entry); });
    if (rc) return rc;
    g->graph_replicated = true;
    g->rows_sharded = false;
    g->n_rows = n;
    return PGVRX_OK;
} catch (const std::bad_alloc &) {
    return gfail(g, PGVRX_ERR_NOMEM, "out of host memory");
}

int pgvrx_group_search_batch(pgvrx_group *g, const void *queries, int64_t nq, int32_t ef_search, int32_t k, int32_t *out_ids,
                             float *out_dists, int32_t *out_counts, pgvrx_group_stats *stats) try {
    if (!g) return PGVRX_ERR_INVALID;
    if (stats) memset(stats, 0, sizeof *stats);
    if (!g->graph_replicated) return gfail(g, PGVRX_ERR_STATE, "no graph loaded (pgvrx_group_load_graph)");
    if (nq < 0 || k < 1 || (nq > 0 && (!queries || !out_ids || !out_dists || !out_counts))) return gfail(g, PGVRX_ERR_INVALID, "bad arguments");
    const int n = (int)g->ix.size();
    const size_t rb = g->ix[0]->row_bytes;
    const double t0 = now_ms();
    // contiguous, balanced query slices: the first nq % n members take one extra query
    int rc = for_members(g, [&](int i) {
        const int64_t base = nq / n, extra = nq % n;
        const int64_t lo = i * base + std::min<int64_t>(i, extra), cnt = base + (i < extra ? 1 : 0);
        if (cnt == 0) return (int)PGVRX_OK;
        return pgvrx_search_batch(g->ix[(size_t)i], (const uint8_t *)queries + (size_t)lo * rb, cnt, ef_search, k, out_ids + (size_t)lo * k,
                                  out_dists + (size_t)lo * k, out_counts + lo, nullptr);
    });
    if (rc) return rc;
    if (stats) { stats->total_ms = now_ms() - t0; stats->n_devices = n; }
    return PGVRX_OK;
} catch (const std::bad_alloc &) {
    return gfail(g, PGVRX_ERR_NOMEM, "out of host memory");
}

// ---- sharded table: exact flat kNN with the merge reading peer memory --------------------------------------------
int pgvrx_group_load_rows(pgvrx_group *g, const void *rows, int64_t n) try {
    if (!g) return PGVRX_ERR_INVALID;
    if (n < 0 || (n > 0 && !rows)) return gfail(g, PGVRX_ERR_INVALID, "bad rows");
    const int nd = (int)g->ix.size();
    const size_t rb = g->ix[0]->row_bytes;
    for (int i = 0; i < nd; i++) {
        const int64_t base = n / nd, extra = n % nd;
        g->lo[(size_t)i] = i * base + std::min<int64_t>(i, extra);
        g->hi[(size_t)i] = g->lo[(size_t)i] + base + (i < extra ? 1 : 0);
    }
    int rc = for_members(g, [&](int i) {
        return pgvrx_index_load_rows(g->ix[(size_t)i], (const uint8_t *)rows + (size_t)g->lo[(size_t)i] * rb, g->hi[(size_t)i] - g->lo[(size_t)i]);
    });
    if (rc) return rc;
    g->rows_sharded = true;
    g->graph_replicated = false;
    g->n_rows = n;
    return PGVRX_OK;
} catch (const std::bad_alloc &) {
    return gfail(g, PGVRX_ERR_NOMEM, "out of host memory");
}

int pgvrx_group_flat_topk(pgvrx_group *g, const void *queries, int64_t nq, int32_t k, int32_t *out_ids, float *out_dists,
                          pgvrx_group_stats *stats) try {
    if (!g) return PGVRX_ERR_INVALID;
    if (stats) memset(stats, 0, sizeof *stats);
    if (!g->rows_sharded) return gfail(g, PGVRX_ERR_STATE, "no rows loaded (pgvrx_group_load_rows)");
    if (nq < 0 || k < 1 || k > FLAT_MAX_K || (nq > 0 && (!queries || !out_ids || !out_dists))) return gfail(g, PGVRX_ERR_INVALID, "bad arguments");
    if (nq == 0) return PGVRX_OK;
    const int nd = (int)g->ix.size();
    const size_t rb = g->ix[0]->row_bytes;
    const double t0 = now_ms();
    // phase 1 (parallel, one host thread per device): queries up, top-k of the device's shard with GLOBAL ids
    int rc = for_members(g, [&](int i) {
        pgvrx_index *ix = g->ix[(size_t)i];
        pgvrx_group::Member &m = g->mem[(size_t)i];
        CU_TRY(ix, cudaSetDevice(g->dev[(size_t)i]));
        CU_TRY(ix, m.q.ensure((size_t)nq * rb));
        CU_TRY(ix, m.ids.ensure((size_t)nq * k * 4));
        CU_TRY(ix, m.dists.ensure((size_t)nq * k * 4));
        CU_TRY(ix, cudaMemcpyAsync(m.q.p, queries, (size_t)nq * rb, cudaMemcpyHostToDevice, ix->stream));
        int r = pgvrx_flat_topk_device(ix, m.q.p, nq, k, 0, -1, g->lo[(size_t)i], (int32_t *)m.ids.p, (float *)m.dists.p, nullptr, ix->stream);
        if (r) return r;
        CU_TRY(ix, cudaEventRecord(g->done[(size_t)i], ix->stream));
        CU_TRY(ix, cudaStreamSynchronize(ix->stream));
        return (int)PGVRX_OK;
    });
    if (rc) return rc;
    const double t1 = now_ms();
    // phase 2 (lead device): one merge kernel over the members' lists -- read in place over NVLink where peer access
    // exists, staged through a peer copy otherwise
    pgvrx_index *lead = g->ix[0];
    G_TRY(g, cudaSetDevice(g->dev[0]));
    G_TRY(g, g->lead_ids.ensure((size_t)nq * k * 4));
    G_TRY(g, g->lead_dists.ensure((size_t)nq * k * 4));
    MergeParts parts;
    parts.n_parts = nd;
    int n_staged = 0;
    for (int i = 0; i < nd; i++) if (!g->peer[(size_t)i]) n_staged++;
    if (n_staged) {
        G_TRY(g, g->stage_ids.ensure((size_t)n_staged * nq * k * 4));
        G_TRY(g, g->stage_dists.ensure((size_t)n_staged * nq * k * 4));
    }
    if (nd > MERGE_MAX_PARTS) return gfail(g, PGVRX_ERR_UNSUPPORTED, "more shards than the merge kernel takes");
    int si = 0;
    for (int i = 0; i < nd; i++) {
        if (g->peer[(size_t)i]) {
            parts.ids[i] = (const int32_t *)g->mem[(size_t)i].ids.p;
            parts.dists[i] = (const float *)g->mem[(size_t)i].dists.p;
        } else {
            int32_t *di = (int32_t *)g->stage_ids.p + (size_t)si * nq * k;
            float *dd = (float *)g->stage_dists.p + (size_t)si * nq * k;
            G_TRY(g, cudaMemcpyPeerAsync(di, g->dev[0], g->mem[(size_t)i].ids.p, g->dev[(size_t)i], (size_t)nq * k * 4, lead->stream));
            G_TRY(g, cudaMemcpyPeerAsync(dd, g->dev[0], g->mem[(size_t)i].dists.p, g->dev[(size_t)i], (size_t)nq * k * 4, lead->stream));
            parts.ids[i] = di;
            parts.dists[i] = dd;
            si++;
        }
    }
    G_TRY(g, cudaEventRecord(g->ev0, lead->stream));
    G_TRY(g, launch_topk_merge_parts(parts, (int)nq, k, k, (int32_t *)g->lead_ids.p, (float *)g->lead_dists.p, lead->stream));
    G_TRY(g, cudaEventRecord(g->ev1, lead->stream));
    G_TRY(g, cudaMemcpyAsync(out_ids, g->lead_ids.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, lead->stream));
    G_TRY(g, cudaMemcpyAsync(out_dists, g->lead_dists.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, lead->stream));
    G_TRY(g, cudaStreamSynchronize(lead->stream));
    if (stats) {
        float mms = 0.0f;
        G_TRY(g, cudaEventElapsedTime(&mms, g->ev0, g->ev1));
        stats->local_ms = t1 - t0;
        stats->merge_kernel_ms = mms;
        stats->total_ms = now_ms() - t0;
        stats->n_devices = nd;
        stats->n_peer_direct = nd - n_staged;
    }
    return PGVRX_OK;
} catch (const std::bad_alloc &) {
    return gfail(g, PGVRX_ERR_NOMEM, "out of host memory");
}

}  // extern "C"
