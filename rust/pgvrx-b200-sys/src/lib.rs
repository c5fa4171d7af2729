//! pgvrx-b200-sys -- raw FFI bindings to libpgvrx_b200.so (the B200-native HNSW hot path for pgvector-rx).
//!
//! GENERATED from include/pgvrx_b200.h by tools/gen_rust_sys.py -- do not edit; re-run the generator instead.
//! The header is the contract (it cites the reference seam each entry point replaces); this file mirrors every
//! constant, enum value, struct field and prototype of it, in header order.  The safe wrapper crate is
//! rust/pgvrx-b200 (with the standalone bench binary), the modules pgvector-rx gains at its two seams
//! (src/index/scan.rs, src/index/build.rs) are rust/pgvector-rx-shim.
#![allow(non_camel_case_types, non_upper_case_globals)]

use std::os::raw::{c_char, c_int, c_void};

pub const PGVRX_ABI_VERSION: i32 = 1;
pub const PGVRX_MIN_M: i32 = 2;
pub const PGVRX_MAX_M: i32 = 100;
pub const PGVRX_DEFAULT_M: i32 = 16;
pub const PGVRX_MIN_EF_CONSTRUCTION: i32 = 4;
pub const PGVRX_MAX_EF_CONSTRUCTION: i32 = 1000;
pub const PGVRX_DEFAULT_EF_CONSTRUCTION: i32 = 64;
pub const PGVRX_MIN_EF_SEARCH: i32 = 1;
pub const PGVRX_MAX_EF_SEARCH: i32 = 1000;
pub const PGVRX_DEFAULT_EF_SEARCH: i32 = 40;
pub const PGVRX_DEFAULT_MAX_SCAN_TUPLES: i32 = 20000;
pub const PGVRX_HEAPTIDS: i32 = 10;
pub const PGVRX_MAX_DIM_VECTOR: i32 = 2000;
pub const PGVRX_MAX_DIM_HALFVEC: i32 = 4000;
pub const PGVRX_MAX_DIM_BIT: i32 = 64000;
pub const PGVRX_ELEM_DELETED: i32 = 1;
pub const PGVRX_ELEM_NO_HEAPTIDS: i32 = 2;

// pgvrx_status
pub const PGVRX_OK: i32 = 0;
pub const PGVRX_ERR_INVALID: i32 = 1;
pub const PGVRX_ERR_CUDA: i32 = 2;
pub const PGVRX_ERR_NOMEM: i32 = 3;
pub const PGVRX_ERR_STATE: i32 = 4;
pub const PGVRX_ERR_UNSUPPORTED: i32 = 5;

// pgvrx_elem
pub const PGVRX_ELEM_F32: i32 = 0;
pub const PGVRX_ELEM_F16: i32 = 1;
pub const PGVRX_ELEM_BIT: i32 = 2;

// pgvrx_metric
pub const PGVRX_METRIC_L2SQ: i32 = 0;
pub const PGVRX_METRIC_NEG_IP: i32 = 1;
pub const PGVRX_METRIC_L1: i32 = 2;
pub const PGVRX_METRIC_HAMMING: i32 = 3;
pub const PGVRX_METRIC_JACCARD: i32 = 4;

// pgvrx_iterative
pub const PGVRX_ITER_OFF: i32 = 0;
pub const PGVRX_ITER_RELAXED: i32 = 1;
pub const PGVRX_ITER_STRICT: i32 = 2;

#[repr(C)]
pub struct PgvrxIndex {
    _private: [u8; 0],
}

#[repr(C)]
pub struct PgvrxScan {
    _private: [u8; 0],
}

#[repr(C)]
pub struct PgvrxGroup {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct PgvrxIndexDesc {
    pub elem_type: i32,
    pub metric: i32,
    pub dim: i32,
    pub m: i32,
    pub ef_construction: i32,
    pub device: i32,
    pub capacity: i64,
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct PgvrxSearchStats {
    pub n_dist: u64,
    pub n_expand0: u64,
    pub n_expand_upper: u64,
    pub n_visited: u64,
    pub n_retried: u64,
    pub n_launches: u32,
    pub kernel_ms: f32,
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct PgvrxPagesInfo {
    pub n_elements: i64,
    pub n_upper_rows: i64,
    pub n_heap_tids: i64,
    pub n_deleted: i64,
    pub entry: i64,
    pub entry_level: i32,
    pub dims: i32,
    pub m: i32,
    pub ef_construction: i32,
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct PgvrxBuildStats {
    pub n_rounds: u64,
    pub n_searched: u64,
    pub n_conflicts: u64,
    pub n_duplicates: u64,
    pub n_big_retries: u64,
    pub n_dist: u64,
    pub n_expand0: u64,
    pub n_expand_upper: u64,
    pub n_pair: u64,
    pub n_pair_gpu: u64,
    pub n_validation_dist: u64,
    pub n_host_dist: u64,
    pub gpu_ms: f64,
    pub kernel_ms: f64,
    pub total_ms: f64,
    pub host_threads: i32,
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct PgvrxVacuumStats {
    pub n_heap_tids_removed: u64,
    pub n_elements_deleted: u64,
    pub n_elements_repaired: u64,
    pub n_rounds: u64,
    pub n_searched: u64,
    pub n_conflicts: u64,
    pub kernel_ms: f64,
    pub total_ms: f64,
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct PgvrxFlatStats {
    pub n_pairs: u64,
    pub n_candidates: u64,
    pub n_uncertified: u64,
    pub n_launches: u32,
    pub n_splits: u32,
    pub path: u32,
    pub kernel_ms: f32,
    pub phase_ms: [f32; 4],
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct PgvrxGroupStats {
    pub total_ms: f64,
    pub local_ms: f64,
    pub merge_kernel_ms: f32,
    pub n_devices: i32,
    pub n_peer_direct: i32,
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct PgvrxGatherProbeResult {
    pub bytes_per_launch: u64,
    pub ms_per_launch: f64,
    pub gb_per_s: f64,
    pub warps_per_sm: i32,
    pub grid: i32,
    pub row_bytes: u32,
}

#[link(name = "pgvrx_b200")]
extern "C" {
    pub fn pgvrx_abi_version() -> c_int;
    pub fn pgvrx_last_error(idx: *const PgvrxIndex) -> *const c_char;
    pub fn pgvrx_device_count() -> c_int;
    pub fn pgvrx_normalize(elem_type: i32, dim: i32, src: *const c_void, out: *mut c_void, norm: *mut f64) -> c_int;
    pub fn pgvrx_normalize_rows(elem_type: i32, dim: i32, src: *const c_void, out: *mut c_void, n: i64, norms: *mut f64) -> c_int;
    pub fn pgvrx_level_from_uniform(r: f64, m: i32) -> i32;
    pub fn pgvrx_max_level(m: i32) -> i32;
    pub fn pgvrx_row_bytes(elem_type: i32, dim: i32) -> usize;
    pub fn pgvrx_index_create(desc: *const PgvrxIndexDesc, out: *mut *mut PgvrxIndex) -> c_int;
    pub fn pgvrx_index_destroy(idx: *mut PgvrxIndex);
    pub fn pgvrx_index_load_graph(idx: *mut PgvrxIndex, rows: *const c_void, n: i64, level: *const u8, adj0: *const i32, upper_base: *const i32, adj_up: *const i32, n_upper_rows: i64, entry: i64) -> c_int;
    pub fn pgvrx_index_size(idx: *const PgvrxIndex) -> i64;
    pub fn pgvrx_pages_inspect(pages: *const c_void, n_pages: i64, elem_type: i32, info: *mut PgvrxPagesInfo) -> c_int;
    pub fn pgvrx_pages_export(pages: *const c_void, n_pages: i64, elem_type: i32, rows: *mut c_void, level: *mut u8, adj0: *mut i32, upper_base: *mut i32, adj_up: *mut i32, heap_tids: *mut u64, heap_tid_counts: *mut u8) -> c_int;
    pub fn pgvrx_pages_write(elem_type: i32, dims: i32, m: i32, ef_construction: i32, n: i64, rows: *const c_void, level: *const u8, adj0: *const i32, upper_base: *const i32, adj_up: *const i32, n_upper_rows: i64, entry: i64, heap_tids: *const u64, heap_tid_counts: *const u8, pages: *mut c_void, cap_pages: i64, n_pages_out: *mut i64) -> c_int;
    pub fn pgvrx_heap_tid_stream(elem_ids: *const i32, n_ids: i64, heap_tids: *const u64, heap_tid_counts: *const u8, n_elements: i64, out_tids: *mut u64, cap: i64, n_out: *mut i64) -> c_int;
    pub fn pgvrx_pages_element_flags(pages: *const c_void, n_pages: i64, flags: *mut u8) -> c_int;
    pub fn pgvrx_index_set_element_flags(idx: *mut PgvrxIndex, flags: *const u8, n: i64) -> c_int;
    pub fn pgvrx_index_build(idx: *mut PgvrxIndex, rows: *const c_void, n_rows: i64, levels: *const i32, heap_tids: *const u64, host_threads: i32, stats: *mut PgvrxBuildStats) -> c_int;
    pub fn pgvrx_index_insert(idx: *mut PgvrxIndex, rows: *const c_void, n_rows: i64, levels: *const i32, heap_tids: *const u64, host_threads: i32, stats: *mut PgvrxBuildStats) -> c_int;
    pub fn pgvrx_search_layer_batch(idx: *mut PgvrxIndex, queries: *const c_void, query_ids: *const i32, nq: i64, ep_ids: *const i32, ep_dists: *const f32, ep_counts: *const i32, ep_pitch: i32, ef: i32, layer: i32, out_ids: *mut i32, out_dists: *mut f32, out_counts: *mut i32, out_expanded: *mut i32, exp_cap: i32, out_n_expanded: *mut i32) -> c_int;
    pub fn pgvrx_commit_insert(idx: *mut PgvrxIndex, row: *const c_void, level: i32, lists: *const i32, n_modified: i32, mod_nodes: *const i32, mod_layers: *const i32, mod_lists: *const i32, new_id_out: *mut i64) -> c_int;
    pub fn pgvrx_index_vacuum(idx: *mut PgvrxIndex, dead_tids: *const u64, n_dead: i64, host_threads: i32, stats: *mut PgvrxVacuumStats) -> c_int;
    pub fn pgvrx_index_element_flags(idx: *mut PgvrxIndex, flags: *mut u8, n: i64) -> c_int;
    pub fn pgvrx_index_set_heap_tids(idx: *mut PgvrxIndex, heap_tids: *const u64, counts: *const u8, n: i64) -> c_int;
    pub fn pgvrx_index_graph_info(idx: *const PgvrxIndex, n: *mut i64, n_upper_rows: *mut i64, entry: *mut i64, entry_level: *mut i32) -> c_int;
    pub fn pgvrx_index_export_graph(idx: *mut PgvrxIndex, rows: *mut c_void, level: *mut u8, adj0: *mut i32, upper_base: *mut i32, adj_up: *mut i32, heap_tids: *mut u64, heap_tid_counts: *mut u8) -> c_int;
    pub fn pgvrx_search_batch(idx: *mut PgvrxIndex, queries: *const c_void, nq: i64, ef_search: i32, k: i32, out_ids: *mut i32, out_dists: *mut f32, out_counts: *mut i32, stats: *mut PgvrxSearchStats) -> c_int;
    pub fn pgvrx_search_batch_device(idx: *mut PgvrxIndex, d_queries: *const c_void, nq: i64, ef_search: i32, k: i32, d_ids: *mut i32, d_dists: *mut f32, d_counts: *mut i32, stats: *mut PgvrxSearchStats, cuda_stream: *mut c_void) -> c_int;
    pub fn pgvrx_index_load_rows(idx: *mut PgvrxIndex, rows: *const c_void, n: i64) -> c_int;
    pub fn pgvrx_flat_topk(idx: *mut PgvrxIndex, queries: *const c_void, nq: i64, k: i32, out_ids: *mut i32, out_dists: *mut f32, stats: *mut PgvrxFlatStats) -> c_int;
    pub fn pgvrx_flat_topk_device(idx: *mut PgvrxIndex, d_queries: *const c_void, nq: i64, k: i32, row_begin: i64, row_end: i64, id_base: i64, d_ids: *mut i32, d_dists: *mut f32, stats: *mut PgvrxFlatStats, cuda_stream: *mut c_void) -> c_int;
    pub fn pgvrx_flat_tc_scores(idx: *mut PgvrxIndex, queries: *const c_void, nq: i64, out_scores: *mut f32, out_eps: *mut f32) -> c_int;
    pub fn pgvrx_rerank(idx: *mut PgvrxIndex, queries: *const c_void, nq: i64, cand_ids: *const i32, nc: i32, k: i32, out_ids: *mut i32, out_dists: *mut f32) -> c_int;
    pub fn pgvrx_rerank_device(idx: *mut PgvrxIndex, d_queries: *const c_void, nq: i64, d_cand: *const i32, nc: i32, k: i32, id_base: i64, d_ids: *mut i32, d_dists: *mut f32, cuda_stream: *mut c_void) -> c_int;
    pub fn pgvrx_topk_merge_device(idx: *mut PgvrxIndex, d_part_ids: *const i32, d_part_dists: *const f32, n_parts: i32, nq: i64, k_in: i32, k_out: i32, d_out_ids: *mut i32, d_out_dists: *mut f32, cuda_stream: *mut c_void) -> c_int;
    pub fn pgvrx_topk_merge_parts_device(idx: *mut PgvrxIndex, d_part_ids: *const *const i32, d_part_dists: *const *const f32, n_parts: i32, nq: i64, k_in: i32, k_out: i32, d_out_ids: *mut i32, d_out_dists: *mut f32, cuda_stream: *mut c_void) -> c_int;
    pub fn pgvrx_group_create(desc: *const PgvrxIndexDesc, devices: *const i32, n_devices: i32, out: *mut *mut PgvrxGroup) -> c_int;
    pub fn pgvrx_group_destroy(g: *mut PgvrxGroup);
    pub fn pgvrx_group_last_error(g: *const PgvrxGroup) -> *const c_char;
    pub fn pgvrx_group_size(g: *const PgvrxGroup) -> i32;
    pub fn pgvrx_group_member(g: *mut PgvrxGroup, i: i32) -> *mut PgvrxIndex;
    pub fn pgvrx_group_peer_access(g: *const PgvrxGroup, i: i32) -> c_int;
    pub fn pgvrx_group_load_graph(g: *mut PgvrxGroup, rows: *const c_void, n: i64, level: *const u8, adj0: *const i32, upper_base: *const i32, adj_up: *const i32, n_upper_rows: i64, entry: i64) -> c_int;
    pub fn pgvrx_group_search_batch(g: *mut PgvrxGroup, queries: *const c_void, nq: i64, ef_search: i32, k: i32, out_ids: *mut i32, out_dists: *mut f32, out_counts: *mut i32, stats: *mut PgvrxGroupStats) -> c_int;
    pub fn pgvrx_group_load_rows(g: *mut PgvrxGroup, rows: *const c_void, n: i64) -> c_int;
    pub fn pgvrx_group_flat_topk(g: *mut PgvrxGroup, queries: *const c_void, nq: i64, k: i32, out_ids: *mut i32, out_dists: *mut f32, stats: *mut PgvrxGroupStats) -> c_int;
    pub fn pgvrx_gather_probe(idx: *mut PgvrxIndex, mode: i32, warps_per_sm: i32, rows_per_step: i32, depth: i32, steps: i32, consume: i32, out: *mut PgvrxGatherProbeResult) -> c_int;
    pub fn pgvrx_scan_begin(idx: *mut PgvrxIndex, queries: *const c_void, nq: i64, ef_search: i32, iterative_mode: i32, max_scan_tuples: i64, out: *mut *mut PgvrxScan) -> c_int;
    pub fn pgvrx_scan_next(scan: *mut PgvrxScan, want_more: *const u8, out_ids: *mut i32, out_dists: *mut f32, out_counts: *mut i32, out_done: *mut u8, stats: *mut PgvrxSearchStats) -> c_int;
    pub fn pgvrx_scan_next_rounds(scan: *mut PgvrxScan, want_more: *const u8, rounds: i32, out_ids: *mut i32, out_dists: *mut f32, out_counts: *mut i32, out_done: *mut u8, stats: *mut PgvrxSearchStats) -> c_int;
    pub fn pgvrx_scan_end(scan: *mut PgvrxScan);
}
