// hnsw_kernels.h -- launch interface between the C-ABI layer (capi.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace pgvrx {

struct GraphView {
    const uint8_t *vectors;     // n rows, row_stride bytes apart, zero padded to 16 B
    const int32_t *adj0;        // n x 2m, -1 after the last neighbour
    const int32_t *adj_up;      // upper rows x m
    const int32_t *upper_base;  // n: first upper row or -1
    const uint8_t *level;       // n
    const uint8_t *flags;       // n or NULL: ELEM_FLAG_* of elements a vacuum touched (NULL for a freshly built index)
    int32_t n, m, dim;
    int32_t entry, entry_level;
    uint32_t row_stride;        // bytes, multiple of 16
};

// per-element state a vacuum leaves on disk (types/hnsw.rs:110-128)
constexpr uint32_t ELEM_FLAG_DELETED = 1;      // HnswElementTupleData.deleted: visited, never evaluated or added (scan.rs:179)
constexpr uint32_t ELEM_FLAG_NO_HEAPTIDS = 2;  // no valid heap TID: searched as usual, emits nothing (scan.rs:866-868)

constexpr int SEARCH_THREADS = 128;   // 4 warps = 4 queries in flight per CTA

struct SearchParams {
    const uint8_t *queries;    // device, q_stride bytes apart, q_bytes valid bytes each
    uint32_t q_stride, q_bytes;
    const int32_t *qmap;       // retry launch: indices of the queries to run (else NULL)
    const int *n_work;         // retry launch: number of entries in qmap (device)
    int nq;
    int ef, k, tail_cap;
    int32_t *out_ids;          // nq x k
    float *out_dists;          // nq x k
    int32_t *out_counts;       // nq
    uint32_t *out_status;      // nq
    uint32_t *out_ctr;         // nq x 4 (n_dist, n_expand0, n_expand_upper, n_visited) or NULL
    uint32_t *out_phase;       // nq x 8 per-phase cycle counts (PGVRX_PHASE_TIMING builds only) or NULL
    uint32_t *vis_base;        // per warp-slot visited tables
    uint32_t vis_entries;      // entries per table, power of two >= 256
    uint32_t up_entries;       // entries of the table the ef = 1 descent uses per upper layer (<= vis_entries): small in the
                               // main pass (it is cleared once per layer), the whole table in the retry pass
    int *work_counter;         // zeroed before the launch
    uint32_t stage_stride;     // bytes between staged rows (row_stride, +16 when that avoids bank conflicts)
    int R;                     // staged rows per warp (<= 32)
    uint32_t stage_bytes;      // staging area per warp
    uint32_t chunk_bytes, chunk_stride;   // wide rows: staged in chunks of chunk_bytes for 32 rows at once (0 = whole rows)
    int spec;                  // 1: gather every neighbour's row before the visited probes resolve (see search_layer); 2: only the first round's
};

// K2: build-side search of a batch of pending inserts (hnsw_build.cu)
struct BuildParams {
    const int32_t *elem_ids;   // element id (= row in the vector table) of each pending insert
    int n_pending;
    int efc, tail_cap;
    int max_layers;            // output rows per insert; layer lc of insert w lives at [w][lc]
    int32_t *w_ids;            // n_pending x max_layers x efc, ascending (distance, id)
    float *w_dists;
    int32_t *w_counts;         // n_pending x max_layers, -1 for layers that were not searched
    int32_t *rs_log;           // n_pending x rs_cap x 3: expanded node id, ordered f_dist bits at the expansion (-1: W not full), layer
    int rs_cap;
    int rs_ent;                // ints per log entry: 3 + ceil(2m / 32) near-mask words
    int32_t *rs_counts;        // n_pending (may exceed rs_cap => set unknown)
    uint32_t *out_status;      // n_pending
    uint32_t *out_ctr;         // n_pending x 4 or NULL
    uint32_t *vis_base;
    uint32_t vis_entries;
    uint32_t up_entries;       // table entries used by the ef = 1 descent (<= vis_entries)
    int *work_counter;
    uint32_t stage_stride;
    int R;
    uint32_t stage_bytes, chunk_bytes, chunk_stride;   // as in SearchParams
};

// K2t: the same search with one CTA per pending insert (hnsw_team.cu): a master warp runs search_layer's sequential rules,
// helper warps compute the neighbour-list distances of the next candidates ahead of it
constexpr int TEAM_MAX_SLOTS = 32;     // cached expansions per CTA (one per lane of the master for lookups)
constexpr int TEAM_MAX_WARPS = 8;
struct TeamGeom {
    int warps;                 // 1 master + (warps - 1) helpers
    int n_slots;               // cache slots (<= TEAM_MAX_SLOTS)
    int lmp;                   // entries per slot: 2m rounded up to 32
    uint32_t vis_entries;      // shared-memory visited table (power of two)
    uint32_t stage_bytes;      // per warp
    uint32_t row_stride;
    int list_j, ef, tail_cap;
};
size_t team_smem_bytes(const TeamGeom &t);
cudaError_t team_kernel_occupancy(int elem, int metric, int list_j, size_t smem_block, int threads, int *blocks_per_sm);
cudaError_t launch_team_build(int elem, int metric, int list_j, const GraphView &g, const BuildParams &p, const TeamGeom &t, int grid,
                              size_t smem_block, cudaStream_t stream);

// vacuum repair searches (hnsw_build.cu: hnsw_repair_search_kernel): find_element_neighbors_on_disk with a skip set
// (src/index/insert.rs:1021-1123 called from vacuum.rs:341-353)
struct RepairParams {
    const int32_t *elem_ids;   // elements to repair (already in the graph)
    int n_pending;
    int ef, tail_cap;          // ef = ef_construction + 1 (insert.rs:1080-1084); list capacity = ef + tail_cap
    int max_layers;
    int w_pitch;               // entries per (element, layer) in w_ids / w_dists: W holds skip members on top of ef
    int32_t *w_ids;            // n_pending x max_layers x w_pitch, ascending (distance, id), skip members included
    float *w_dists;
    int32_t *w_counts;         // n_pending x max_layers, -1 for layers that were not searched
    int32_t *rs_log;           // expansion log as in BuildParams
    int rs_cap, rs_ent;
    int32_t *rs_counts;
    uint32_t *out_status;
    uint32_t *out_ctr;
    uint32_t *vis_base;
    uint32_t vis_entries;
    uint32_t up_entries;       // table entries used by the ef = 1 descent (<= vis_entries)
    int *work_counter;
    uint32_t stage_stride;
    int R;
    uint32_t stage_bytes, chunk_bytes, chunk_stride;   // as in SearchParams
};
cudaError_t launch_repair_search(int elem, int metric, const GraphView &g, const RepairParams &p, int grid, size_t smem_block,
                                 int threads, cudaStream_t stream);
cudaError_t repair_kernel_occupancy(int elem, int metric, size_t smem_block, int threads, int *blocks_per_sm);

// check_element_closer masks for groups of <= PAIR_MAX_ROWS elements (hnsw_build.cu)
constexpr int PAIR_MAX_ROWS = 128;
struct PairParams {
    const int32_t *ids;        // concatenated element ids of all groups
    const float *own_dist;     // per member: its distance to the group's node
    const int32_t *group_off;  // n_groups + 1 offsets into ids, or NULL in W mode:
    const int32_t *w_counts;   //   W mode: group gi has w_counts[gi] members (<= 0: none) at ids + gi * w_pitch,
    int w_pitch, w_layers;     //   gi = pending * w_layers + layer
    int n_groups;
    uint32_t *out;             // per member PAIR_MAX_ROWS / 32 words: bit b set <=> d(member, b) <= own_dist[member]
    int *work_counter;
    uint32_t stage_stride;
};

// what the host needs to size shared memory
struct SearchLaunch {
    int R, ef, tail_cap;
    int list_j;                // 0: result list in shared memory; 2/4/8: registers per lane
    uint32_t stage_stride, row_stride;
    uint32_t stage_bytes;      // staging area per warp: R whole rows, or 32 chunks of a wide row
    uint32_t chunk_bytes, chunk_stride;
};

size_t search_warp_smem(const SearchLaunch &L);
cudaError_t scan_kernel_occupancy(int elem, int metric, int list_j, bool flagged, size_t smem_block, int threads, int *blocks_per_sm);
cudaError_t launch_scan_kernel(int elem, int metric, int list_j, const GraphView &g, const SearchParams &p, int grid,
                               size_t smem_block, int threads, cudaStream_t stream);
cudaError_t build_kernel_occupancy(int elem, int metric, int list_j, bool flagged, size_t smem_block, int threads, int *blocks_per_sm);
cudaError_t launch_build_search(int elem, int metric, int list_j, const GraphView &g, const BuildParams &p, int grid,
                                size_t smem_block, int threads, cudaStream_t stream);
cudaError_t launch_pairwise(int elem, int metric, const GraphView &g, const PairParams &p, int max_rows, int grid, cudaStream_t stream);
cudaError_t launch_pending_pairs(int elem, int metric, const GraphView &g, const int32_t *ids, int n, float *out, int sm_count,
                                 cudaStream_t stream);
cudaError_t launch_scatter_rows(int32_t *dst, const int64_t *row_offsets, const int32_t *packed, int row_len, int n_rows,
                                cudaStream_t stream);
cudaError_t launch_collect_failed(const uint32_t *status, int nq, int32_t *qmap, int *n_work, cudaStream_t stream);


// ---- K4 / K6 / re-rank (flat_knn.cu) ----------------------------------------------------------------
constexpr int FLAT_THREADS = 128;   // one base row per thread per tile
constexpr int FLAT_QT = 32;         // queries per CTA: one accumulator chain per (thread, query)
constexpr int FLAT_CB = 512;        // bytes of a row staged per dimension chunk
constexpr int FLAT_MAX_K = 1024;

struct FlatGeom {
    uint32_t chunk_bytes, rstride, q_chunk_bytes;
    uint32_t rows_bytes, qs_bytes, list_bytes;
    int n_chunks;
};

struct FlatParams {
    const uint8_t *vectors;    // row 0 of the table (local row numbering)
    uint32_t row_stride;
    int64_t row_begin, row_end;   // rows [row_begin, row_end) are scanned
    int64_t rows_per_split;       // multiple of FLAT_THREADS; blockIdx.y selects the split
    int64_t id_base;              // reported id = id_base + (row - row_begin)
    const uint8_t *queries;
    uint32_t q_stride, q_bytes;
    int nq, k;
    int32_t *out_ids;          // [n_splits][nq][k], ascending (distance, id), -1 / +inf padded
    float *out_dists;
    const uint8_t *flags;      // per table row or NULL: rows with a non-zero flag (deleted / no heap TID left) are skipped
};

struct RerankParams {
    const uint8_t *vectors;
    uint32_t row_stride, stage_stride;
    int64_t n_rows, id_base;   // candidate ids are id_base + local row
    const uint8_t *queries;
    uint32_t q_stride, q_bytes;
    const int32_t *cand;       // nq x nc, -1 = no candidate
    int nq, nc, k, R;
    int32_t *out_ids;          // nq x k
    float *out_dists;
    const uint8_t *flags;      // per table row or NULL: flagged candidates are dropped
};

size_t flat_smem_bytes(int elem, uint32_t row_stride, int k, FlatGeom *out);
cudaError_t flat_kernel_occupancy(int elem, int metric, size_t smem, int *blocks_per_sm);
cudaError_t launch_flat_partial(int elem, int metric, const FlatParams &p, const FlatGeom &g, int n_qtiles, int n_splits,
                                size_t smem, cudaStream_t stream);
cudaError_t launch_topk_merge(const int32_t *ids, const float *dists, int n_parts, int nq, int k_in, int k_out,
                              int32_t *out_ids, float *out_dists, cudaStream_t stream);
// the same merge over per-shard lists that live in separate allocations -- other GPUs' memory included: with peer
// access enabled the kernel's loads go over NVLink, so the gather and the merge are one pass (group.cu)
constexpr int MERGE_MAX_PARTS = 64;
struct MergeParts {
    int n_parts;
    const int32_t *ids[MERGE_MAX_PARTS];      // each [nq][k_in]
    const float *dists[MERGE_MAX_PARTS];
};
cudaError_t launch_topk_merge_parts(const MergeParts &parts, int nq, int k_in, int k_out, int32_t *out_ids, float *out_dists,
                                    cudaStream_t stream);
size_t rerank_smem_bytes(uint32_t row_stride, uint32_t stage_stride, int R, int k);
cudaError_t launch_rerank(int elem, int metric, const RerankParams &p, int grid, size_t smem, cudaStream_t stream);


// ---- K3: iterative scan (hnsw_iter.cu) ------------------------------------------------------------------
constexpr int ITER_CHUNK = 64;            // keys per chunk of a query's discarded set
constexpr int ITER_EXT = 32;              // chunks per extent (consecutive chunk ids: one warp-wide load reads their minima)
constexpr int ITER_UP_ENTRIES = 4096;     // per-warp hashed visited table for the ef = 1 descent
constexpr int ITER_TAIL_CAP = 1024;       // retained boundary ties behind W (bit types)
constexpr int ITER_TAIL_CAP_FLOAT = 256;  // float types
constexpr int ITER_WARPS = 1;             // warps (= queries in flight) per CTA: finest occupancy granularity

struct IterQueryState {
    long long tuples;          // scan.rs:871
    uint32_t prev_ord;         // ordered bits of previous_distance (scan.rs:801-806); 0 = -infinity
    int round;                 // rounds this query has run
    int head, tail, fill, live;   // discarded: first extent / chunk taking appends (-1: none), keys in that chunk, live keys
    int done;
    int pad;
};

struct IterParams {
    const uint8_t *queries;
    uint32_t q_stride, q_bytes;
    int nq, ef, tail_cap;
    int mode;                  // pgvrx_iterative
    int rounds;                // refills per launch (>= 1); outputs are nq x rounds x ef / nq x rounds
    long long max_scan_tuples;
    const uint8_t *want_more;  // nq or NULL
    IterQueryState *state;
    uint32_t *bitmap;          // nq x n_words visited bits
    uint32_t n_words;
    uint64_t *pool;            // discarded keys: extents x ITER_EXT chunks x ITER_CHUNK
    uint64_t *cmin;            // per chunk: smallest key
    int *cfill;                // per chunk: keys held
    int *next;                 // per extent: next extent of the owning query's chain
    int *pool_next;            // extents handed out
    int pool_cap;              // extents
    uint32_t *vis_up;          // per warp slot, ITER_UP_ENTRIES
    int32_t *out_ids;          // nq x ef
    float *out_dists;
    int32_t *out_counts;
    uint8_t *out_done;
    uint32_t *out_status;
    uint32_t *out_ctr;         // nq x 4 or NULL
    int *work_counter;
    uint32_t stage_stride;
    int R;
};

size_t iter_warp_smem(int R, uint32_t stage_stride, uint32_t row_stride, int ef, int tail_cap);
cudaError_t iter_kernel_occupancy(int elem, int metric, size_t smem_block, int *blocks_per_sm);
cudaError_t launch_iter_scan(int elem, int metric, const GraphView &g, const IterParams &p, int grid, size_t smem_block,
                             cudaStream_t stream);


// ---- K4b: 3xTF32 tcgen05 candidate generation (flat_tc.cu) ------------------------------------------------
constexpr int TC_M = 128;          // queries per tile (UMMA M)
constexpr int TC_N = 128;          // base rows per tile (UMMA N)
constexpr int TC_KB = 32;          // fp32 elements per K block = one 128-byte swizzle row
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 192;    // TMA warp, MMA warp, four epilogue warps
constexpr int TC_KP = 32;          // candidates kept per (query, split)
constexpr int TC_MAX_K = 24;       // largest k served by the tensor-core path

struct TcParams {
    int nq, n_mtiles;
    int64_t n_rows;                // valid base rows (the operand arrays are padded to n_ntiles * TC_N)
    int n_ntiles, n_splits, ntiles_per_split;
    int d_pad, kp, metric;
    int a_resident;                // d_pad <= 128: the query tile stays in shared memory for the whole unit
    int64_t id_base;
    const float *qnorm;            // |q|^2, padded to n_mtiles * TC_M
    const float *xnorm;            // |x|^2, padded
    int32_t *cand_ids;             // [nq][n_splits][kp] global ids, -1 padded
    float *cand_bound;             // [n_splits][nq]: approximate distance every unkept row of the split is at least
    float *dbg_scores;             // nq x n_rows approximate distances (diagnostic) or NULL
    uint32_t *hint;                // nq: ordered bits of the smallest full candidate-set maximum seen so far per query
    int tile_step;                 // a unit visits tiles nt0, nt0 + tile_step, ... (1 = all of them; > 1: the sample pre-pass)
    int kp_eff;                    // keys a candidate set really holds (TC_KP, or fewer in a set-maintaining pre-pass)
    int local_k;                   // 8 / 16: hint pre-pass with lane-local top lists (no sets); 0: sets or append mode
    // append mode (after a pre-pass): no candidate sets -- every row whose approximate distance is <= hint[q] is appended to
    // query q's list; the certificate's bound is the hint, a list that overflows leaves its query uncertified
    int32_t *app_ids;              // nq x n_splits x app_cap global ids (caller pre-fills -1) or NULL: one segment per unit
    int *app_cnt;                  // nq x n_splits: rows the unit wanted to append (> app_cap: some were dropped)
    int app_cap;
};

size_t tc_smem_bytes(int kp);
cudaError_t launch_tc_split(const uint8_t *src, uint32_t src_stride, int64_t n_rows, int64_t n_pad, int dim, int d_pad, float *hi,
                            float *lo, float *norm, uint32_t *norm_max, uint32_t *hint_init, cudaStream_t stream);
cudaError_t launch_flat_tc(const void *map_qhi, const void *map_qlo, const void *map_xhi, const void *map_xlo, const TcParams &p,
                           int grid, cudaStream_t stream);
cudaError_t launch_flat_tc2(const void *map_qhi, const void *map_qlo, const void *map_xhi, const void *map_xlo, const TcParams &p,
                            int grid, cudaStream_t stream);
cudaError_t launch_tc_certify(const float *topk_dists, const int32_t *topk_ids, int k, const float *cand_bound, int n_splits, int nq,
                              const float *qnorm, const uint32_t *xnorm_max, float eps_rel, int metric, const uint32_t *hint,
                              const int *app_cnt, int app_cap, int app_splits, int32_t *uncert_list, int *n_uncert, cudaStream_t stream);
cudaError_t launch_tc_gather_queries(const uint8_t *q, uint32_t q_stride, const int32_t *list, int n, uint8_t *out, cudaStream_t stream);
cudaError_t launch_tc_scatter_results(const int32_t *ids, const float *dists, int k, const int32_t *list, int n, int32_t *out_ids,
                                      float *out_dists, cudaStream_t stream);

}  // namespace pgvrx
