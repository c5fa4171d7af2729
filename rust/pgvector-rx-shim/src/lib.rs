//! pgvector-rx GPU shim -- the two seams of SURVEY section 8(b) as Rust modules.
//!
//! In pgvector-rx this directory becomes `src/gpu/` (`mod gpu;` in src/lib.rs).  Nothing here touches PostgreSQL types
//! directly: the AM callbacks keep detoasting, normalisation, locking, WAL and page writes, and hand plain byte slices
//! and ids across.  The call-site changes in the reference are listed at the top of each module with the lines they
//! replace; INTEGRATION.md sections 3 and 4 walk through them.
//!
//! * `gpu_scan`  : `get_scan_items` / `resume_scan_items` (src/index/scan.rs:458-577) -> `ScanBridge`
//! * `gpu_build` : `build_callback`'s insert (src/index/build.rs:400-535) -> `BuildBridge`; `aminsert`'s neighbour search
//!                 and back-link rule (src/index/insert.rs:1021-1123, 500-738) -> `BuildBridge::insert_batch`
//! * `guc`       : the GUC / reloption values the bridges take, with the reference's names, ranges and defaults
pub mod gpu_build;
pub mod gpu_scan;
pub mod guc;
