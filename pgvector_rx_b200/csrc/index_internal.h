// index_internal.h -- handle layout and small helpers shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <memory>
#include <string>

#include "../../include/pgvrx_b200.h"
#include "hnsw_kernels.h"

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grow while keeping the first `keep` bytes (device-to-device copy)
    cudaError_t ensure_keep(size_t bytes, size_t keep) {
        if (bytes <= cap) return cudaSuccess;
        void *np = nullptr;
        size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) return e;
        if (p && keep) e = cudaMemcpy(np, p, keep, cudaMemcpyDeviceToDevice);
        if (p) cudaFree(p);
        p = np;
        cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

inline uint32_t round16u(size_t x) { return (uint32_t)((x + 15) & ~(size_t)15); }

namespace pgvrx_host {
void set_create_error(const std::string &msg);
struct BuildState;                       // host mirror of the graph during/after pgvrx_index_build
void destroy_build_state(BuildState *);
}

struct pgvrx_index {
    pgvrx_index_desc desc{};
    size_t row_bytes = 0;
    uint32_t row_stride = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_ph[4] = {nullptr, nullptr, nullptr, nullptr};   // phase boundaries inside one flat-kNN call (stats only)
    // graph in HBM
    DevBuf vectors, adj0, adj_up, upper_base, level;
    DevBuf flags;               // per-element vacuum state (pgvrx_index_set_element_flags); used only when has_flags
    bool has_flags = false;
    int64_t n = 0, n_upper = 0, entry = -1;
    int entry_level = -1;
    // bumped by everything that changes the graph or the table (load_graph, load_rows, build, insert, vacuum,
    // set_element_flags): an open pgvrx_scan remembers the value it was begun under and refuses to continue on another
    uint64_t generation = 0;
    bool loaded = false;        // graph + rows resident
    bool rows_loaded = false;   // vector table resident (enough for the flat / re-rank paths)
    // scratch
    DevBuf q, ids, dists, counts;
    // Scan scratch (visited tables, per-query status, work counters).  One set per stream the caller scans on, so that
    // batches submitted on different streams overlap on the device (the tail of one persistent launch is filled by the
    // head of the next); a set is rebound to a new stream only after its previous stream has drained.
    struct ScanScratch {
        DevBuf status, ctr, vis, vis_retry, qmap, work, phase;
        cudaStream_t owner = nullptr;
        bool used = false;
    };
    static constexpr int N_SCAN_SCRATCH = 4;
    ScanScratch scan_scratch[N_SCAN_SCRATCH];
    int scan_scratch_rr = 0;
    DevBuf part_ids, part_dists, cand;   // flat kNN partial lists, re-rank candidates
    // tensor-core flat path (K4b): tf32 hi/lo splits of the table (built lazily, dropped when the table changes)
    DevBuf tc_xhi, tc_xlo, tc_xnorm, tc_xmax, tc_qhi, tc_qlo, tc_qnorm, tc_cand, tc_bound, tc_uncert, tc_cnt, tc_fq, tc_fids,
        tc_fdists, tc_dbg, tc_hint, tc_appcnt;
    bool tc_ready = false;
    int64_t tc_n_pad = 0;
    int tc_d_pad = 0;
    alignas(64) unsigned char tc_map_xhi[128], tc_map_xlo[128];
    pgvrx_host::BuildState *build = nullptr;
    std::string err;
};

int pgvrx_fail(pgvrx_index *idx, int code, const std::string &msg);

#define CU_TRY(idx, expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            char _b[512];                                                                       \
            snprintf(_b, sizeof _b, "%s: %s (%s:%d)", cudaGetErrorName(_e), cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                       \
            cudaGetLastError();                                                                 \
            return pgvrx_fail(idx, _e == cudaErrorMemoryAllocation ? PGVRX_ERR_NOMEM : PGVRX_ERR_CUDA, _b); \
        }                                                                                       \
    } while (0)

pgvrx::GraphView pgvrx_view_of(const pgvrx_index *ix);
uint32_t pgvrx_pow2_at_least(uint64_t x);

// one search configuration: staging geometry, scratch sizes, grid
struct ScanConfig {
    pgvrx::SearchLaunch L;
    uint32_t vis_entries;
    size_t smem_block;
    int grid;
    int wpb;                   // warps (= queries in flight) per CTA
};
// tail_cap < 0: pick the register-resident list when ef is small enough.  for_build selects the K2
// kernel's occupancy instead of K1's.
int pgvrx_make_config(pgvrx_index *ix, int ef, int tail_cap, uint32_t vis_entries, int64_t nq, int max_grid,
                      bool for_build, ScanConfig *out);
