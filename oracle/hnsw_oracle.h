/*
 * hnsw_oracle.h -- CPU restatement of pgvector-rx's HNSW hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (pgvector_rx_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED against every PG-free known answer the reference's own tests hold
 * for this path (see tests/test_oracle_known_answers.py: graph/mod.rs:538-778 unit tests,
 * vector.rs/halfvec.rs/bitvec.rs scalar distance answers, pg_regress expected orders); the aminsert and
 * ambulkdelete restatements by the reference's insert pg_tests and its TAP recall thresholds for the insert and
 * vacuum paths (tests/test_oracle_live_index.py: tests/t/011, 013, 014, 022 on seeded data).
 * UNPINNED (third-party code absent from /root/reference):
 *   - tie order inside Rust std::collections::BinaryHeap (restated from its published
 *     algorithm as ORC_TIE_RUST_HEAP; the reference has no test that pins a tie order);
 *   - rand::random::<f64>() levels (unseeded in the reference): levels[] is an input here.
 *
 * All citations are file:line under /root/reference.
 */
#ifndef HNSW_ORACLE_H
#define HNSW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element storage type (opclass input type: vector / halfvec / bit) */
enum { ORC_ELEM_F32 = 0, ORC_ELEM_F16 = 1, ORC_ELEM_BIT = 2 };
/* opclass FUNCTION 1 (src/types/vector.rs:842-856, halfvec.rs:1050-1063, bitvec.rs:210-237).
 * cosine opclasses = NEG_IP over rows normalised by FUNCTION 2 (orc_normalize_*). */
enum { ORC_METRIC_L2SQ = 0, ORC_METRIC_NEG_IP = 1, ORC_METRIC_L1 = 2,
       ORC_METRIC_HAMMING = 3, ORC_METRIC_JACCARD = 4 };
/* tie policy for equal distances (SURVEY 8a note T) */
enum { ORC_TIE_RUST_HEAP = 0,  /* std BinaryHeap sift order, distance-only Ord        */
       ORC_TIE_TOTAL_ORDER = 1 /* (distance, element id) lexicographic -- what the GPU does */ };
/* hnsw.iterative_scan (src/index/options.rs:85,168-175) */
enum { ORC_ITER_OFF = 0, ORC_ITER_RELAXED = 1, ORC_ITER_STRICT = 2 };

typedef struct orc_index orc_index;

typedef struct {
    uint64_t n_dist;         /* distance evaluations (graph/mod.rs:218, scan.rs:375)        */
    uint64_t n_expand0;      /* candidates expanded at layer 0 (graph/mod.rs:203, scan.rs:349) */
    uint64_t n_expand_upper; /* candidates expanded at layers >= 1                           */
    uint64_t n_visited;      /* ids inserted into visited sets                               */
    uint64_t n_pair;         /* pairwise distances inside select_neighbors (graph/mod.rs:326) */
} orc_counters;

/* ---- scalar kernels (a13-a16) ------------------------------------------------------- */
float    orc_half_to_f32(uint16_t h);                 /* halfvec.rs:54-87  */
uint16_t orc_f32_to_half(float f);                    /* halfvec.rs:92-143 */
/* distance as the opclass support function returns it (f64). a/b point at raw element data:
 * float[dim], uint16_t[dim] or ceil(dim/8) bytes. */
double orc_distance(int elem_type, int metric, int dim, const void *a, const void *b);
/* FUNCTION 2 + normalise (vector.rs:106-126,672-683; halfvec.rs:204-233,832-843).
 * returns the f64 norm; writes the normalised row to out unless norm == 0 (out zeroed).
 * returns -1.0 on halfvec overflow (reference raises "value out of range: overflow"). */
double orc_normalize(int elem_type, int dim, const void *in, void *out);
/* build.rs:373-377 with r supplied by the caller; max level per types/hnsw.rs:337-349 */
int orc_level_from_uniform(double r, int m);
int orc_max_level(int m);

/* ---- index lifecycle ------------------------------------------------------------------ */
orc_index *orc_create(int elem_type, int metric, int dim, int m, int ef_construction, int tie_mode);
void orc_destroy(orc_index *);
size_t orc_row_bytes(const orc_index *);

/* a5: the ambuild insert loop (build.rs:400-535) over n rows in order.  rows are raw element
 * data, row_bytes apart, ALREADY normalised when the opclass is cosine (caller does a16 and
 * drops zero-norm rows, as build.rs:426-438 does).  levels[i] is the level of the i-th row.
 * heap_tid of row i is base_tid + i.  Returns number of graph elements after the call. */
int64_t orc_build(orc_index *, const void *rows, int64_t n, const int32_t *levels, int64_t base_tid);

/* aminsert (src/index/insert.rs:1227-1480) for n rows in order, on whatever graph the index holds: nearest-lm lists
 * for the new element (no heuristic, insert.rs:1096-1117), one-slot back-link updates chosen by get_update_index
 * (insert.rs:500-738), duplicate merge (insert.rs:1136-1213), entry point on a strictly greater level.  Page
 * placement and slot reuse are not restated: new elements get the next id.  Returns the element count. */
int64_t orc_insert(orc_index *, const void *rows, int64_t n, const int32_t *levels, int64_t base_tid);

/* graph introspection (element ids are insertion order among kept elements) */
int64_t orc_n_elements(const orc_index *);
int64_t orc_entry_point(const orc_index *);        /* -1 when empty */
int     orc_element_level(const orc_index *, int64_t id);
/* copies up to cap neighbour ids (and distances if dists != NULL) of (id, layer) in stored
 * order; returns the list length */
int     orc_neighbors(const orc_index *, int64_t id, int layer, int32_t *ids, float *dists, int cap);
int     orc_heap_tids(const orc_index *, int64_t id, int64_t *tids, int cap);
const void *orc_element_row(const orc_index *, int64_t id);
void    orc_build_counters(const orc_index *, orc_counters *out);

/* dense export used to feed the GPU library the same graph:
 *   level[n] (u8), adj0[n * 2m] (-1 padded), upper_base[n] (-1 when level 0), and the upper
 *   rows adj_up[n_upper_rows * m] where (id, layer l>=1) lives at row upper_base[id] + l - 1.
 * Call with NULL outputs to size: returns n_upper_rows. */
int64_t orc_export_graph(const orc_index *, uint8_t *level, int32_t *adj0, int32_t *upper_base,
                         int32_t *adj_up);
/* load a graph (inverse of export) without building; neighbour distances are recomputed */
int orc_import_graph(orc_index *, const void *rows, int64_t n, const uint8_t *level,
                     const int32_t *adj0, const int32_t *upper_base, const int32_t *adj_up,
                     int64_t entry);

/* ---- what a vacuum leaves on disk (scan-side semantics only: scan.rs:179,262,866-868) --------- */
enum { ORC_VACUUM_REMOVE_HEAPTIDS = 1, /* vacuum.rs:118-190: element keeps its place in the graph, emits nothing */
       ORC_VACUUM_MARK_DELETED = 2,    /* vacuum.rs:717-758: deleted = 1, lists and value cleared, version bumped */
       ORC_VACUUM_STALE_NEIGHBORS = 3  /* neighbour tuple whose version no longer matches (scan.rs:262)            */ };
enum { ORC_ELEM_DELETED = 1, ORC_ELEM_NO_HEAPTIDS = 2 };
/* ambulkdelete (src/index/vacuum.rs:816-842) with the heap's dead TIDs given as a list: removes them, repairs the
 * lists of the live elements that pointed at emptied elements (skip-set searches, vacuum.rs:288-407), marks the
 * emptied elements deleted.  Returns the number of heap TIDs removed. */
int64_t orc_bulk_delete(orc_index *, const int64_t *dead_tids, int64_t n_dead);
int  orc_vacuum_mark(orc_index *, int64_t id, int what);
void orc_element_flags(const orc_index *, uint8_t *flags /* n */);

/* ---- a1-a4 as free-standing calls on the index's graph (unit-test surface) -------------- */
/* search_layer (graph/mod.rs:161-255). entry points are (ids, dists); returns count written,
 * ascending by distance. query is raw element data. */
int orc_search_layer(orc_index *, const void *query, const int32_t *ep_ids, const float *ep_dists,
                     int n_ep, int ef, int layer, int32_t *out_ids, float *out_dists);
/* select_neighbors (graph/mod.rs:269-308) */
int orc_select_neighbors(orc_index *, const int32_t *cand_ids, const float *cand_dists, int n_cand,
                         int max_neighbors, int32_t *out_ids, float *out_dists);

/* ---- scan (a8-a12) ---------------------------------------------------------------------- */
/* get_scan_items + amgettuple emission, iterative_scan = off (scan.rs:458-530,794-875).
 * For each query: up to k elements nearest first.  out_ids/out_dists are nq*k, padded with
 * -1 / +inf.  queries must be normalised already for cosine.  nthreads > 1 partitions
 * queries statically (CPU baseline only; results are per-query independent). */
int orc_search_batch(orc_index *, const void *queries, int64_t nq, int ef_search, int k,
                     int32_t *out_ids, float *out_dists, int32_t *out_counts,
                     orc_counters *counters /* summed over queries, may be NULL */, int nthreads);

/* Full amgettuple stream for one query with iterative scan semantics (scan.rs:709-875):
 * emits heap TIDs in the order the executor would receive them until the scan is exhausted
 * or max_emit tuples were produced.  Returns number emitted.  elem_ids/dists (optional)
 * give the element each TID came from. */
int64_t orc_scan_stream(orc_index *, const void *query, int ef_search, int iterative_mode,
                        int64_t max_scan_tuples, int64_t max_emit, int64_t *out_tids,
                        int32_t *out_elem_ids, double *out_dists, orc_counters *counters);

/* ---- a20: exact kNN (sequential scan + sort by operator value) -------------------------- */
int orc_flat_topk(int elem_type, int metric, int dim, const void *rows, int64_t n,
                  const void *queries, int64_t nq, int k, int32_t *out_ids, float *out_dists,
                  int nthreads);

/* ---- a19: on-disk pages (test fixture writer; build.rs:545-821, types/hnsw.rs:18-169) ----- */
/* serialises the index exactly as the flush phase does; returns number of 8 KiB pages.
 * Call with pages == NULL to size.  heap TID t is written as (block t >> 16, offset (t & 0xffff) + 1). */
int64_t orc_write_pages(const orc_index *, uint8_t *pages, int64_t cap_pages);

#ifdef __cplusplus
}
#endif
#endif
