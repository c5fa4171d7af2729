/*
 * pgvrx_b200.h -- C ABI of the B200-native HNSW hot path for pgvector-rx.
 *
 * This is the drop-in boundary: the entry points a pgrx shim (or any FFI) binds in place of
 * the reference's in-process graph code.  Plain pointers and sizes only; the library owns all
 * device memory behind the opaque handle; callers own every host buffer; nothing here throws,
 * unwinds or longjmps -- every call returns a pgvrx_status and leaves a message retrievable with
 * pgvrx_last_error() (the reference raises pgrx::error!, e.g. src/index/scan.rs:136,733).
 *
 * Reference seams replaced (file:line under /root/reference):
 *   scan : get_scan_items / resume_scan_items          src/index/scan.rs:458-577
 *          (search_layer_disk + load_element + load_neighbor_tids, scan.rs:155-448)
 *   build: find_element_neighbors' search_layer calls  src/graph/mod.rs:355-427 (:386,:407)
 *          distance_fn                                  src/index/build.rs:346-369
 *   exact: operator-ordered sequential scan (ground truth of tests/t/012..., no dedicated fn)
 * Parameter names, ranges and defaults mirror the reloptions and GUCs in
 * src/index/options.rs:81-96,203-225 and src/hnsw_constants.rs:47-112.
 *
 * There is no CPU fallback: every compute entry point fails with PGVRX_ERR_CUDA when no
 * sm_100-class device is usable.
 */
#ifndef PGVRX_B200_H
#define PGVRX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGVRX_ABI_VERSION 1

typedef enum {
    PGVRX_OK = 0,
    PGVRX_ERR_INVALID = 1,     /* bad argument / option out of the reference's range            */
    PGVRX_ERR_CUDA = 2,        /* CUDA runtime or device error (message has the cudaError name) */
    PGVRX_ERR_NOMEM = 3,
    PGVRX_ERR_STATE = 4,       /* call order (e.g. search before a graph is loaded)             */
    PGVRX_ERR_UNSUPPORTED = 5
} pgvrx_status;

/* indexed type: `vector`, `halfvec`, `bit` (src/types/{vector,halfvec,bitvec}.rs) */
typedef enum { PGVRX_ELEM_F32 = 0, PGVRX_ELEM_F16 = 1, PGVRX_ELEM_BIT = 2 } pgvrx_elem;

/* opclass FUNCTION 1.  vector_l2_ops/halfvec_l2_ops -> L2SQ (squared, vector.rs:842-845);
 * *_ip_ops -> NEG_IP; *_cosine_ops -> NEG_IP over rows/queries normalised by the caller with
 * pgvrx_normalize (FUNCTION 2, vector.rs:852-856, build.rs:426-438, scan.rs:749-754);
 * *_l1_ops -> L1; bit_hamming_ops / bit_jaccard_ops (bitvec.rs:210-237). */
typedef enum {
    PGVRX_METRIC_L2SQ = 0, PGVRX_METRIC_NEG_IP = 1, PGVRX_METRIC_L1 = 2,
    PGVRX_METRIC_HAMMING = 3, PGVRX_METRIC_JACCARD = 4
} pgvrx_metric;

/* hnsw.iterative_scan (options.rs:85,168-175) */
typedef enum { PGVRX_ITER_OFF = 0, PGVRX_ITER_RELAXED = 1, PGVRX_ITER_STRICT = 2 } pgvrx_iterative;

/* reloptions / GUC ranges (hnsw_constants.rs:47-112) */
#define PGVRX_MIN_M 2
#define PGVRX_MAX_M 100
#define PGVRX_DEFAULT_M 16
#define PGVRX_MIN_EF_CONSTRUCTION 4
#define PGVRX_MAX_EF_CONSTRUCTION 1000
#define PGVRX_DEFAULT_EF_CONSTRUCTION 64
#define PGVRX_MIN_EF_SEARCH 1
#define PGVRX_MAX_EF_SEARCH 1000
#define PGVRX_DEFAULT_EF_SEARCH 40
#define PGVRX_DEFAULT_MAX_SCAN_TUPLES 20000
#define PGVRX_HEAPTIDS 10
#define PGVRX_MAX_DIM_VECTOR 2000      /* hnsw_constants.rs:4     */
#define PGVRX_MAX_DIM_HALFVEC 4000     /* halfvec.rs:875-877      */
#define PGVRX_MAX_DIM_BIT 64000        /* bitvec.rs:183-185       */

typedef struct pgvrx_index pgvrx_index;

typedef struct {
    int32_t elem_type;        /* pgvrx_elem                                                  */
    int32_t metric;           /* pgvrx_metric                                                */
    int32_t dim;              /* dimensions (bits for PGVRX_ELEM_BIT)                        */
    int32_t m;                /* reloption m, 2..100                                         */
    int32_t ef_construction;  /* reloption ef_construction, 4..1000 and >= 2*m               */
    int32_t device;           /* CUDA device ordinal                                          */
    int64_t capacity;         /* max graph elements the handle must hold                      */
} pgvrx_index_desc;

/* per-call work counters: the algorithmic-bytes figure of DESIGN.md is derived from these */
typedef struct {
    uint64_t n_dist;          /* distance evaluations (scan.rs:375, graph/mod.rs:218)          */
    uint64_t n_expand0;       /* candidates expanded at layer 0 (scan.rs:349)                  */
    uint64_t n_expand_upper;  /* candidates expanded at layers >= 1                            */
    uint64_t n_visited;       /* ids added to visited sets                                     */
    uint64_t n_retried;       /* queries re-run with larger scratch (visited/tie overflow)     */
    uint32_t n_launches;      /* kernels launched by this call                                 */
    float    kernel_ms;       /* device time of those kernels (CUDA events on the call stream) */
} pgvrx_search_stats;

/* ---- library ---------------------------------------------------------------------------- */
int pgvrx_abi_version(void);
/* message for the last failure on this handle (or the last failed create when idx == NULL);
 * valid until the next call on the same handle / thread. */
const char *pgvrx_last_error(const pgvrx_index *idx);
/* number of usable CUDA devices (0 when none: every compute call then fails) */
int pgvrx_device_count(void);

/* ---- host-side opclass support (bit-identical to the reference; no device needed) ---------- */
/* FUNCTION 2 + normalisation (vector.rs:106-126, halfvec.rs:204-233).  Writes the unit row to
 * out (may alias in).  *norm receives the f64 norm; rows with norm 0 must not be indexed
 * (build.rs:433-435).  PGVRX_ERR_INVALID on halfvec overflow ("value out of range: overflow"). */
int pgvrx_normalize(int32_t elem_type, int32_t dim, const void *in, void *out, double *norm);
/* the same for n consecutive rows (norms may be NULL); a row with norm 0 comes out as a zero row */
int pgvrx_normalize_rows(int32_t elem_type, int32_t dim, const void *in, void *out, int64_t n, double *norms);
/* build.rs:373-377: floor(-ln(max(r, MIN_POSITIVE)) / ln m) capped at the page-size limit
 * (types/hnsw.rs:337-349).  r is the caller's uniform(0,1) draw. */
int32_t pgvrx_level_from_uniform(double r, int32_t m);
int32_t pgvrx_max_level(int32_t m);
/* bytes per stored row: dim*4, dim*2 or ceil(dim/8) */
size_t pgvrx_row_bytes(int32_t elem_type, int32_t dim);

/* ---- index lifecycle ------------------------------------------------------------------------ */
int pgvrx_index_create(const pgvrx_index_desc *desc, pgvrx_index **out);
void pgvrx_index_destroy(pgvrx_index *idx);

/* Upload a finished graph (from the page exporter, or any builder) into HBM.
 *   rows        n x row_bytes raw element data, element id == row number
 *   level       n x u8                                     (HnswElementTupleData.level)
 *   adj0        n x 2m int32 element ids, -1 after the last neighbour (layer 0)
 *   upper_base  n x int32: first row in adj_up of an element with level >= 1, else -1
 *   adj_up      n_upper_rows x m int32; (id, layer l >= 1) is row upper_base[id] + l - 1
 *   entry       entry point element id (HnswMetaPageData.entry_*), -1 for an empty index
 * Neighbour order is the stored order: it is the order search visits them (scan.rs:273-279). */
int pgvrx_index_load_graph(pgvrx_index *idx, const void *rows, int64_t n, const uint8_t *level,
                           const int32_t *adj0, const int32_t *upper_base, const int32_t *adj_up,
                           int64_t n_upper_rows, int64_t entry);
int64_t pgvrx_index_size(const pgvrx_index *idx);

/* ---- page exporter: HNSW relation pages -> the arrays pgvrx_index_load_graph takes ------------- */
/* `pages` is the index relation's main fork, n_pages * 8192 bytes (block 0 = meta page).  Layout:
 * src/types/hnsw.rs:18-169; walk: src/index/scan.rs:126-283.  Element ids are assigned in page
 * order (= build insertion order, build.rs:600-702).  Host code, no device needed. */
typedef struct {
    int64_t n_elements;
    int64_t n_upper_rows;     /* sum of element levels = rows of adj_up                          */
    int64_t n_heap_tids;
    int64_t n_deleted;        /* elements flagged deleted by a vacuum (exported with a zero row) */
    int64_t entry;            /* element id of HnswMetaPageData.entry_*, -1 when the index is empty */
    int32_t entry_level;
    int32_t dims, m, ef_construction;
} pgvrx_pages_info;
int pgvrx_pages_inspect(const void *pages, int64_t n_pages, int32_t elem_type, pgvrx_pages_info *info);
/* rows: n x row_bytes; level: n; adj0: n x 2m; upper_base: n; adj_up: n_upper_rows x m;
 * heap_tids (optional): n x 10 as (block << 16 | offset), ~0 when absent; heap_tid_counts (optional): n */
int pgvrx_pages_export(const void *pages, int64_t n_pages, int32_t elem_type, void *rows, uint8_t *level,
                       int32_t *adj0, int32_t *upper_base, int32_t *adj_up, uint64_t *heap_tids,
                       uint8_t *heap_tid_counts);

/* ---- page writer: the arrays of pgvrx_index_export_graph -> the index relation's pages ------------------------- */
/* The flush phase of ambuild (src/index/build.rs:545-821: create_meta_page, create_graph_pages,
 * write_neighbor_tuples, update_meta_page) for a graph built by pgvrx_index_build: block 0 = meta page, element and
 * neighbour tuples in element order from block 1, byte for byte what the reference writes from its in-memory build
 * state -- so build + export + write = CREATE INDEX.  heap_tids / heap_tid_counts as pgvrx_index_export_graph returns
 * them, each TID as (block << 16 | offset) with offset >= 1; NULL gives element i the synthetic TID of row i
 * (block i / 65535, offset i % 65535 + 1: always a valid TID).  Call with pages == NULL to get the page count in *n_pages_out.
 * Host code, no device needed. */
int pgvrx_pages_write(int32_t elem_type, int32_t dims, int32_t m, int32_t ef_construction, int64_t n,
                      const void *rows, const uint8_t *level, const int32_t *adj0, const int32_t *upper_base,
                      const int32_t *adj_up, int64_t n_upper_rows, int64_t entry, const uint64_t *heap_tids,
                      const uint8_t *heap_tid_counts, void *pages, int64_t cap_pages, int64_t *n_pages_out);

/* amgettuple's fan-out (src/index/scan.rs:794-811,858-875): the scan entry points return ELEMENT ids in emission order;
 * every element stands for up to 10 heap tuples, which amgettuple hands to the executor one per call, taking them from
 * the BACK of the element's list (scan.rs:798).  Given n_ids element ids (-1 = padding, skipped) and the heap TID
 * arrays of pgvrx_pages_export / pgvrx_index_export_graph (n_elements x 10 and n_elements counts), writes the heap TIDs
 * in the order xs_heaptid would take them.  *n_out receives the number of TIDs; when it exceeds cap only the first cap
 * were written.  An element without heap TIDs contributes nothing (scan.rs:866-868).  Host-only. */
int pgvrx_heap_tid_stream(const int32_t *elem_ids, int64_t n_ids, const uint64_t *heap_tids, const uint8_t *heap_tid_counts,
                          int64_t n_elements, uint64_t *out_tids, int64_t cap, int64_t *n_out);

/* ---- what a vacuum leaves on disk, as the scan sees it ------------------------------------------ */
/* Per-element flags of an index that vacuum has touched (src/index/vacuum.rs:118-190,647-790):
 *   PGVRX_ELEM_DELETED      HnswElementTupleData.deleted != 0 -- load_element rejects the element before its
 *                           distance is computed (src/index/scan.rs:179): it is marked visited and nothing else;
 *                           a deleted entry point makes the scan return no rows (scan.rs:475-486)
 *   PGVRX_ELEM_NO_HEAPTIDS  no valid heap TID left -- searched like any element, but amgettuple skips it and it
 *                           does not count towards hnsw.max_scan_tuples (scan.rs:866-871)
 * A neighbour tuple whose version or count no longer matches its element (scan.rs:262) exports as empty lists.
 * pgvrx_pages_element_flags fills flags[n_elements] from the pages; pgvrx_index_set_element_flags hands them to a
 * loaded index (NULL clears; all-zero flags keep the flag-free kernels).  The in-memory build path has no such
 * elements and refuses to extend an index that carries flags.  The flags are an index-scan matter: the exact paths
 * (pgvrx_flat_topk, pgvrx_rerank) stand for a scan of the TABLE and read whatever rows the caller loaded. */
#define PGVRX_ELEM_DELETED 1
#define PGVRX_ELEM_NO_HEAPTIDS 2
int pgvrx_pages_element_flags(const void *pages, int64_t n_pages, uint8_t *flags);
int pgvrx_index_set_element_flags(pgvrx_index *idx, const uint8_t *flags, int64_t n);

/* ---- build: the ambuild insert loop (src/index/build.rs:400-535) ------------------------------- */
typedef struct {
    uint64_t n_rounds;        /* GPU search rounds                                               */
    uint64_t n_searched;      /* element searches run on the GPU (>= elements: conflicts re-run)  */
    uint64_t n_conflicts;     /* speculative results discarded because an earlier commit of the
                                 same round modified a list the search had read                   */
    uint64_t n_duplicates;    /* rows merged into an existing element's heap TIDs (build.rs:482-512) */
    uint64_t n_big_retries;   /* searches re-run alone with large scratch                         */
    uint64_t n_dist, n_expand0, n_expand_upper;   /* GPU work of the committed searches           */
    uint64_t n_pair;          /* pairwise distances select_neighbors consulted (graph/mod.rs:326)          */
    uint64_t n_pair_gpu;      /* pairwise distances computed on the GPU for back-link re-selections        */
    uint64_t n_validation_dist; /* pending-vs-pending distances consulted when validating speculative searches   */
    uint64_t n_host_dist;     /* distances evaluated on the HOST (0 unless a candidate group does not fit on a CTA:
                                 ef_construction > 128 or very wide rows)                                   */
    double   gpu_ms;          /* wall time spent in upload + kernel + download                     */
    double   kernel_ms;       /* device time of the K2 launches alone (CUDA events)               */
    double   total_ms;
    int32_t  host_threads;
} pgvrx_build_stats;

/* Insert n_rows rows in order, exactly as build_callback does for each heap tuple (NULL rows and,
 * for cosine opclasses, zero-norm rows must be dropped and the rest normalised by the caller, as
 * build.rs:411-438 does before the insert).  levels[i] is the level of row i (the reference draws it
 * from an unseeded RNG, build.rs:373-377; use pgvrx_level_from_uniform).  heap_tids[i] (optional,
 * default i) is recorded per element, up to 10 per element for byte-identical duplicates.  The library treats
 * them as opaque 64-bit values; pass real ItemPointers as (block << 16 | offset), offset >= 1, when the graph is to be
 * written out with pgvrx_pages_write (which refuses offset 0: a reader would take it for "no heap TID").
 * The search_layer calls of find_element_neighbors (graph/mod.rs:386,407) run on the GPU in
 * speculative batches; select_neighbors, update_neighbor_connections, the duplicate rule and the
 * entry-point rule run on the host in the reference's order, and a batched result is only used when
 * it provably equals the sequential one -- the adjacency is identical to a sequential build.
 * May be called again to append more rows.  The graph is resident and searchable afterwards. */
int pgvrx_index_build(pgvrx_index *idx, const void *rows, int64_t n_rows, const int32_t *levels,
                      const uint64_t *heap_tids, int32_t host_threads, pgvrx_build_stats *stats);
/* ---- inserts into a live index: aminsert (src/index/insert.rs:1227-1480) ------------------------- */
/* Same inputs and statistics as pgvrx_index_build, different rules -- the ones the reference applies to an index
 * that already exists on disk: the new element's lists are the nearest lm candidates of each layer's search, without
 * the heuristic (find_element_neighbors_on_disk, insert.rs:1021-1123); a neighbour with a free slot gets the
 * back-link appended, a full list has ONE slot overwritten, chosen by get_update_index (insert.rs:500-738: closer-test
 * heuristic in which the new element is never pruned and never prunes; the first slot whose element was not kept);
 * duplicates merge into an existing element with 1..9 heap TIDs (insert.rs:1136-1213); the entry point moves on a
 * strictly greater level.  Works on a handle built here or loaded with pgvrx_index_load_graph (the host mirror is
 * then rebuilt from the device arrays; loaded elements count as holding one heap TID = their id).  Page placement and
 * the reuse of deleted slots (insert.rs:104-470) stay with the storage engine: new elements get the next id.  Not
 * supported on an index that carries vacuum flags. */
int pgvrx_index_insert(pgvrx_index *idx, const void *rows, int64_t n_rows, const int32_t *levels,
                       const uint64_t *heap_tids, int32_t host_threads, pgvrx_build_stats *stats);

/* ---- the narrow build seam: search_layer on the GPU, heuristic and commit on the CALLER's side --------------------- */
/* SURVEY 8(b)(4)/(5).  A caller that keeps select_neighbors / update_neighbor_connections / the insert loop itself (the
 * pgrx build_callback, src/index/build.rs:400-535) replaces only find_element_neighbors' search_layer calls
 * (src/graph/mod.rs:386,407; signature of the reference's seam: DistanceFn, graph/mod.rs:144-145) and tells the library
 * what it decided.  pgvrx_index_build / pgvrx_index_insert are the same machinery with both halves inside the library.
 *
 * pgvrx_search_layer_batch: search_layer(query, entry_points, ef, layer) (graph/mod.rs:161-255) for nq queries against
 * the resident graph.  Query i is raw element data (queries, row_bytes apart) or the stored row of element query_ids[i]
 * (exactly one of the two is non-NULL).  Its entry points are ep_counts[i] <= min(ep_pitch, ef) (id, distance) pairs at
 * ep_ids / ep_dists + i * ep_pitch, each on `layer`; all of them start out visited, in C and in W (graph/mod.rs:171-184).
 * out_ids / out_dists: nq x ef, W ascending by (distance, id), -1 / +inf padded; out_counts[i] = |W|.  out_expanded
 * (optional, exp_cap ids per query, -1 padded) receives the elements the search expanded, in expansion order -- its read
 * set: a speculative result stays valid while none of those lists has changed; out_n_expanded[i] is the full count (the
 * list is a prefix when it exceeds exp_cap).  HOST buffers. */
int pgvrx_search_layer_batch(pgvrx_index *idx, const void *queries, const int32_t *query_ids, int64_t nq,
                             const int32_t *ep_ids, const float *ep_dists, const int32_t *ep_counts, int32_t ep_pitch,
                             int32_t ef, int32_t layer, int32_t *out_ids, float *out_dists, int32_t *out_counts,
                             int32_t *out_expanded, int32_t exp_cap, int32_t *out_n_expanded);
/* pgvrx_commit_insert: ONE new element as the caller's host code decided it -- appended as the next element id
 * (returned in *new_id_out): `row` (row_bytes), `level`, `lists` = its neighbour lists, layer 0 first (2m ids) then
 * layers 1 .. level (m ids each), -1 after the last neighbour of a list (graph/mod.rs:419-422) -- and the n_modified
 * neighbour lists update_neighbor_connections rewrote (graph/mod.rs:442-489): list t belongs to (mod_nodes[t],
 * mod_layers[t]) and is the next 2m (layer 0) or m ids of mod_lists.  The entry point moves to the new element when its
 * level is strictly greater (build.rs:523-529).  The next search sees the commit.  The handle's host mirror is dropped
 * (pgvrx_index_export_graph / pgvrx_index_insert rebuild it from the device arrays when needed). */
int pgvrx_commit_insert(pgvrx_index *idx, const void *row, int32_t level, const int32_t *lists, int32_t n_modified,
                        const int32_t *mod_nodes, const int32_t *mod_layers, const int32_t *mod_lists, int64_t *new_id_out);

/* ---- VACUUM: ambulkdelete (src/index/vacuum.rs:816-842) ----------------------------------------- */
typedef struct {
    uint64_t n_heap_tids_removed;  /* pass 1 (vacuum.rs:118-216)                                              */
    uint64_t n_elements_deleted;   /* pass 3: elements left without heap TIDs, now flagged deleted (vacuum.rs:655-790) */
    uint64_t n_elements_repaired;  /* pass 2: elements whose lists were re-found (vacuum.rs:288-407)           */
    uint64_t n_rounds;             /* GPU search rounds                                                       */
    uint64_t n_searched;           /* repair searches run (>= repaired: conflicts re-run)                      */
    uint64_t n_conflicts;          /* speculative searches discarded: they had expanded an element repaired
                                      earlier in the same round                                               */
    double   kernel_ms, total_ms;
} pgvrx_vacuum_stats;
/* dead_tids: the heap TIDs the table's vacuum reports dead, in the same (block << 16 | offset) form the index holds.
 * Removes them, re-finds the neighbour lists of every live element that pointed at an emptied element or whose layer-0
 * list is not full -- searches with the skip set, ef_construction + 1 and search_layer_disk's skip_count bookkeeping
 * (scan.rs:314-336,416-429), run on the GPU in batches and committed in element order -- replaces a deleted entry point
 * by the highest remaining element, and marks the emptied elements deleted.  The result equals the reference's
 * sequential vacuum (asserted against the oracle).  Afterwards the index carries element flags
 * (pgvrx_index_element_flags reads them back) and scans apply them; pgvrx_index_export_graph returns the repaired
 * lists.  Needs the heap TIDs: a handle built here, or loaded + pgvrx_index_set_heap_tids. */
int pgvrx_index_vacuum(pgvrx_index *idx, const uint64_t *dead_tids, int64_t n_dead, int32_t host_threads,
                       pgvrx_vacuum_stats *stats);
int pgvrx_index_element_flags(pgvrx_index *idx, uint8_t *flags, int64_t n);

/* Heap TIDs of a loaded index (n x 10 as block << 16 | offset with ~0 padding, and n counts -- the arrays
 * pgvrx_pages_export fills): lets pgvrx_index_insert apply the duplicate rule's capacity test (an element takes a
 * duplicate's TID only while it holds 1..9, insert.rs:1159-1163) and pgvrx_index_export_graph return them.  Without
 * this call every loaded element counts as holding one heap TID equal to its id. */
int pgvrx_index_set_heap_tids(pgvrx_index *idx, const uint64_t *heap_tids, const uint8_t *counts, int64_t n);

int pgvrx_index_graph_info(const pgvrx_index *idx, int64_t *n, int64_t *n_upper_rows, int64_t *entry,
                           int32_t *entry_level);
/* copies the graph out (any pointer may be NULL); shapes as in pgvrx_index_load_graph, heap_tids n x 10 (~0 when
 * absent), heap_tid_counts n.  A handle without a host mirror (loaded, or fed by pgvrx_commit_insert) reads the arrays
 * back from the device first; its elements then count as holding one heap TID equal to their id. */
int pgvrx_index_export_graph(pgvrx_index *idx, void *rows, uint8_t *level, int32_t *adj0,
                             int32_t *upper_base, int32_t *adj_up, uint64_t *heap_tids,
                             uint8_t *heap_tid_counts);

/* ---- scan: get_scan_items, iterative_scan = off (scan.rs:458-530) ------------------------------ */
/* For each of nq queries (raw element data, row_bytes apart, already normalised for cosine):
 * runs the greedy descent (ef = 1 per upper layer, fresh visited set per layer) and the layer-0
 * search with ef_search, and returns the k <= ef_search nearest elements nearest-first
 * (amgettuple pops results nearest-first, scan.rs:861).  out_ids / out_dists are nq*k, padded
 * with -1 / +inf; out_counts[q] = number of valid entries.  Equal distances are ordered by
 * element id.  HOST buffers; the call copies in, launches, copies out and synchronises. */
int pgvrx_search_batch(pgvrx_index *idx, const void *queries, int64_t nq, int32_t ef_search,
                       int32_t k, int32_t *out_ids, float *out_dists, int32_t *out_counts,
                       pgvrx_search_stats *stats /* may be NULL */);
/* Same, with DEVICE-accessible buffers on `cuda_stream` (a cudaStream_t; NULL = the handle's stream): device memory
 * or page-locked host memory (the kernels then read queries from / write results to the host directly).
 * Asynchronous unless stats != NULL (stats needs the kernels to have finished); calls on different streams overlap on
 * the device (one scratch set per stream).  A query that exhausted its scratch even in the retry pass is reported
 * with out_counts[q] = -1 (the blocking pgvrx_search_batch turns that into an error). */
int pgvrx_search_batch_device(pgvrx_index *idx, const void *d_queries, int64_t nq, int32_t ef_search,
                              int32_t k, int32_t *d_ids, float *d_dists, int32_t *d_counts,
                              pgvrx_search_stats *stats, void *cuda_stream);

/* ---- exact paths: flat kNN, re-rank, top-k merge ---------------------------------------------- */
/* The reference has no flat-kNN routine: its ground truth is PostgreSQL's sequential scan ordered by
 * the distance operator (tests/t/012_hnsw_vector_build_recall.pl:83, 013...pl:95-98; operators
 * src/types/vector.rs:586-659, halfvec.rs:765-817, bitvec.rs:144-167).  These entry points are that
 * scan on the GPU: every distance is the opclass FUNCTION 1 value computed with the reference's
 * arithmetic (bit-identical), results ordered by (distance, element id). */
typedef struct {
    uint64_t n_pairs;         /* (query, row) distance evaluations                                */
    uint64_t n_candidates;    /* rows re-ranked exactly (tensor-core path only)                   */
    uint64_t n_uncertified;   /* queries the error bound could not certify, re-run on the exact path */
    uint32_t n_launches;
    uint32_t n_splits;        /* base-row splits merged per query                                 */
    uint32_t path;            /* 0 = sequential f32 chains on CUDA cores; 1 = 3xTF32 tcgen05 candidate
                                 generation + exact re-rank (large fp32 L2 / IP batches);
                                 2 = the same on CTA pairs (cta_group::2)                         */
    float    kernel_ms;       /* device time of the call's kernels (CUDA events)                  */
    float    phase_ms[4];     /* tensor-core path only, device time per phase: [0] operand split of the
                                 queries (+ of the table on its first use), [1] candidate kernel (GEMM +
                                 per-query candidate sets), [2] exact re-rank, [3] certification and the
                                 exact re-run of uncertified queries; zeros on the CUDA-core path    */
} pgvrx_flat_stats;

/* Upload a vector table without a graph (enough for the exact paths).  Element id == row number. */
int pgvrx_index_load_rows(pgvrx_index *idx, const void *rows, int64_t n);
/* k nearest rows per query by exact distance.  out_ids / out_dists are nq*k, padded with -1 / +inf
 * when the table has fewer than k rows.  HOST buffers. */
int pgvrx_flat_topk(pgvrx_index *idx, const void *queries, int64_t nq, int32_t k, int32_t *out_ids,
                    float *out_dists, pgvrx_flat_stats *stats /* may be NULL */);
/* DEVICE buffers; scans rows [row_begin, row_end) (row_end < 0: to the end) and reports
 * id_base + (row - row_begin): a rank that holds one shard of a larger base set passes the shard's
 * first global id.  Asynchronous on cuda_stream unless stats != NULL. */
int pgvrx_flat_topk_device(pgvrx_index *idx, const void *d_queries, int64_t nq, int32_t k, int64_t row_begin,
                           int64_t row_end, int64_t id_base, int32_t *d_ids, float *d_dists,
                           pgvrx_flat_stats *stats, void *cuda_stream);
/* Diagnostic of the tensor-core path (fp32 L2 / inner product): the approximate distances it ranks
 * candidates by (nq x n, HOST) and, per query, the bound eps it certifies against
 * (|approximate - exact| <= eps is what makes the final result exact). */
int pgvrx_flat_tc_scores(pgvrx_index *idx, const void *queries, int64_t nq, float *out_scores, float *out_eps);
/* Exact re-rank: cand_ids is nq x nc element ids (-1 = none); returns the k nearest of each list by
 * exact distance, ordered by (distance, id). */
int pgvrx_rerank(pgvrx_index *idx, const void *queries, int64_t nq, const int32_t *cand_ids, int32_t nc,
                 int32_t k, int32_t *out_ids, float *out_dists);
int pgvrx_rerank_device(pgvrx_index *idx, const void *d_queries, int64_t nq, const int32_t *d_cand, int32_t nc,
                        int32_t k, int64_t id_base, int32_t *d_ids, float *d_dists, void *cuda_stream);
/* K-way merge of per-shard results laid out [part][query][k_in] (what an all-gather of every rank's
 * pgvrx_flat_topk_device output produces) into the k_out nearest per query.  DEVICE buffers. */
int pgvrx_topk_merge_device(pgvrx_index *idx, const int32_t *d_part_ids, const float *d_part_dists,
                            int32_t n_parts, int64_t nq, int32_t k_in, int32_t k_out, int32_t *d_out_ids,
                            float *d_out_dists, void *cuda_stream);
/* The same merge over per-part DEVICE pointers (HOST arrays of n_parts <= 64 pointers, each list [query][k_in]): for
 * lists that are not contiguous with each other -- one packed all-gather of (ids | distances) per rank, or peers' lists
 * read in place. */
int pgvrx_topk_merge_parts_device(pgvrx_index *idx, const int32_t *const *d_part_ids, const float *const *d_part_dists,
                                  int32_t n_parts, int64_t nq, int32_t k_in, int32_t k_out, int32_t *d_out_ids,
                                  float *d_out_dists, void *cuda_stream);

/* ---- several GPUs of one box behind one handle (one process; one host thread per device while a call runs) ------- */
/* The scan shards by QUERIES against a graph replicated on every device (no exchange step); the exact flat kNN shards
 * the BASE ROWS contiguously, every device computes the top-k of its shard for all queries with global ids, and the
 * lead device (devices[0]) merges the per-shard lists with one kernel that reads the other devices' lists in place
 * over NVLink (peer access; staged through a peer copy where that is unavailable).  desc->device is ignored: member i
 * runs on devices[i] (an ordinal may repeat: several members then share a device).  One process per GPU launchers use
 * pgvrx_flat_topk_device + an all-gather + pgvrx_topk_merge_device instead (pgvector_rx_b200/sharded.py). */
typedef struct pgvrx_group pgvrx_group;
typedef struct {
    double  total_ms;         /* wall time of the call                                                        */
    double  local_ms;         /* flat: wall time of the per-device phase (queries up + top-k of the shard)     */
    float   merge_kernel_ms;  /* flat: device time of the merge kernel on the lead device                      */
    int32_t n_devices;
    int32_t n_peer_direct;    /* flat: members whose lists the merge read in place (peer access or same device) */
} pgvrx_group_stats;
int pgvrx_group_create(const pgvrx_index_desc *desc, const int32_t *devices, int32_t n_devices, pgvrx_group **out);
void pgvrx_group_destroy(pgvrx_group *g);
const char *pgvrx_group_last_error(const pgvrx_group *g);
int32_t pgvrx_group_size(const pgvrx_group *g);
pgvrx_index *pgvrx_group_member(pgvrx_group *g, int32_t i);          /* the member's own handle (owned by the group)  */
int pgvrx_group_peer_access(const pgvrx_group *g, int32_t i);         /* 1: the lead device reads member i in place    */
/* replicate a finished graph on every member (arguments as pgvrx_index_load_graph) */
int pgvrx_group_load_graph(pgvrx_group *g, const void *rows, int64_t n, const uint8_t *level, const int32_t *adj0,
                           const int32_t *upper_base, const int32_t *adj_up, int64_t n_upper_rows, int64_t entry);
/* pgvrx_search_batch over all members: one caller batch split into contiguous query slices, results written into the
 * caller's arrays (HOST buffers; page-locked ones are read / written by the kernels in place) */
int pgvrx_group_search_batch(pgvrx_group *g, const void *queries, int64_t nq, int32_t ef_search, int32_t k, int32_t *out_ids,
                             float *out_dists, int32_t *out_counts, pgvrx_group_stats *stats /* may be NULL */);
/* shard n rows contiguously over the members (member i holds rows [i*n/N ...), ids stay global row numbers) */
int pgvrx_group_load_rows(pgvrx_group *g, const void *rows, int64_t n);
/* exact k nearest rows per query over the sharded table; results identical to pgvrx_flat_topk on the whole table */
int pgvrx_group_flat_topk(pgvrx_group *g, const void *queries, int64_t nq, int32_t k, int32_t *out_ids, float *out_dists,
                          pgvrx_group_stats *stats /* may be NULL */);

/* ---- diagnostics: the ceiling the scan kernel's row gathers are measured against ------------------------------ */
/* Random row_stride-byte gathers from the handle's vector table with the scan kernel's own transfer mechanism (one TMA
 * bulk copy per row into a per-warp staging area, persistent one-warp CTAs) and nothing else -- no visited set, no
 * distances.  mode 0: row ids uniform over the table; mode 1: a random walk over the layer-0 adjacency (the hub reuse,
 * hence the L2 hit rate, of real searches; needs a loaded graph).  warps_per_sm 0 = as many as fit; rows_per_step rows
 * per mbarrier phase (the scan uses 32 for 512-byte rows), depth phases in flight per warp, consume != 0 reads every
 * staged byte back from shared memory (as the distance loop does).  Measures gathered bytes per second: the denominator
 * bench.py reports next to the HBM copy peak. */
typedef struct {
    uint64_t bytes_per_launch;
    double   ms_per_launch;
    double   gb_per_s;
    int32_t  warps_per_sm, grid;
    uint32_t row_bytes;
} pgvrx_gather_probe_result;
int pgvrx_gather_probe(pgvrx_index *idx, int32_t mode, int32_t warps_per_sm, int32_t rows_per_step, int32_t depth,
                       int32_t steps, int32_t consume, pgvrx_gather_probe_result *out);

/* ---- scan: iterative index scan (hnsw.iterative_scan = relaxed_order | strict_order) ------------- */
/* A pgvrx_scan is what amrescan creates for a batch of queries: per query the visited set, the
 * discarded heap, the tuple count and the strict-order watermark (HnswScanState, scan.rs:586-617)
 * live in HBM behind it.  Each pgvrx_scan_next is, for every query, the block of amgettuple that
 * refills `results` -- get_scan_items on the first call (scan.rs:458-530), then resume_scan_items
 * (scan.rs:538-577), or the one-at-a-time drain of `discarded` once hnsw.max_scan_tuples elements
 * were produced (scan.rs:831-841) -- followed by amgettuple's emission rules: nearest first, each
 * element counts towards `tuples` (scan.rs:871), strict_order drops elements nearer than the last
 * one emitted (scan.rs:801-806).  The caller maps element ids to heap TIDs (last to first,
 * scan.rs:798) and stops asking for a query once its LIMIT is met. */
typedef struct pgvrx_scan pgvrx_scan;
int pgvrx_scan_begin(pgvrx_index *idx, const void *queries, int64_t nq, int32_t ef_search,
                     int32_t iterative_mode /* pgvrx_iterative */, int64_t max_scan_tuples, pgvrx_scan **out);
/* want_more: nq bytes or NULL (= all); a query with 0 is left untouched this round.
 * out_ids / out_dists: nq * ef_search, emitted elements first, -1 / +inf padded; out_counts[q] = number
 * emitted this round; out_done[q] = 1 once the scan of q is exhausted (later rounds emit nothing).
 * PGVRX_ITER_OFF emits get_scan_items' batch and is then done.  HOST buffers. */
int pgvrx_scan_next(pgvrx_scan *scan, const uint8_t *want_more, int32_t *out_ids, float *out_dists,
                    int32_t *out_counts, uint8_t *out_done, pgvrx_search_stats *stats /* may be NULL */);
/* `rounds` refills in one call (one kernel launch, one copy back): amgettuple asks for the next batch only once the
 * previous one is used up (scan.rs:817), so several batches handed out together are the same stream -- a consumer
 * that stops early has merely had the later ones computed for nothing.  Worth it under a selective filter (WHERE i % 500
 * = 0 needs ~125 refills for 10 rows at ef_search = 40).  out_ids / out_dists: nq * rounds * ef_search, batch r of query q
 * at (q * rounds + r) * ef_search; out_counts: nq * rounds; out_done: nq.  pgvrx_scan_next is rounds = 1. */
int pgvrx_scan_next_rounds(pgvrx_scan *scan, const uint8_t *want_more, int32_t rounds, int32_t *out_ids, float *out_dists,
                           int32_t *out_counts, uint8_t *out_done, pgvrx_search_stats *stats /* may be NULL */);
void pgvrx_scan_end(pgvrx_scan *scan);

#ifdef __cplusplus
}
#endif
#endif
