//! Build seam.
//!
//! Call-site change in src/index/build.rs: `build_callback` (build.rs:400-535) keeps the NULL skip, detoast, zero-norm
//! skip and normalisation (:411-438) and the level draw (:457), then calls `BuildBridge::push(row, level, heap_tid)`
//! instead of `find_element_neighbors` + duplicate merge + `update_neighbor_connections` (:464-525).  After
//! `table_index_build_scan` returns (build.rs:873), `ambuild` calls `BuildBridge::finish()` and gets the relation's
//! pages laid out as its own flush phase (build.rs:545-821) lays them out -- byte for byte -- to copy through the
//! buffer manager and `log_newpage_range`.
//!
//! `aminsert` (src/index/insert.rs:1227-1480) keeps detoast / normalise / level / HNSW_UPDATE_LOCK / page placement /
//! WAL; `BuildBridge::insert_batch` decides which lists change (find_element_neighbors_on_disk :1021-1123,
//! find_duplicate_on_disk :1180-1213, get_update_index :500-738) for a batch of rows, equal to one-at-a-time inserts.
//!
//! A maintainer who wants to keep `select_neighbors` / `update_neighbor_connections` in Rust uses the narrow seam
//! instead: `search_layer_batch` + `commit_insert` (graph/mod.rs:144-145,386,407), wrapped at the bottom of this file.
use pgvrx_b200::{level_from_uniform, Elem, GraphArrays, Index, Metric, Result};
use pgvrx_b200_sys as ffi;
use std::os::raw::c_void;

pub struct BuildBridge {
    index: Index,
    rows: Vec<u8>,
    levels: Vec<i32>,
    heap_tids: Vec<u64>,
    row_bytes: usize,
}

/// Encodes an ItemPointer (block, offset) the way the library's heap-TID arrays hold it: block << 16 | offset.
pub fn tid_u64(block: u32, offset: u16) -> u64 {
    ((block as u64) << 16) | offset as u64
}

impl BuildBridge {
    /// `capacity`: reltuples estimate; the library grows its device arrays once to this size.
    pub fn new(elem: Elem, metric: Metric, dim: i32, m: i32, ef_construction: i32, device: i32, capacity: i64) -> Result<BuildBridge> {
        let index = Index::create(elem, metric, dim, m, ef_construction, device, capacity)?;
        let row_bytes = index.row_bytes();
        Ok(BuildBridge { index, rows: Vec::new(), levels: Vec::new(), heap_tids: Vec::new(), row_bytes })
    }

    /// build.rs:373-377 with the caller's uniform draw (`rand::random::<f64>()` in the reference).
    pub fn draw_level(&self, r: f64) -> i32 {
        level_from_uniform(r, self.index.m)
    }

    /// One heap tuple in heap-scan order (= insert order = element order of the finished index).
    pub fn push(&mut self, row: &[u8], level: i32, heap_tid: u64) {
        debug_assert_eq!(row.len(), self.row_bytes);
        self.rows.extend_from_slice(row);
        self.levels.push(level);
        self.heap_tids.push(heap_tid);
    }

    /// Runs the insert loop for everything pushed so far and returns the graph in the arrays the flush phase writes
    /// (element order = insert order, duplicates merged into <= 10 heap TIDs, neighbour lists in the reference's order).
    pub fn finish(&mut self) -> Result<(GraphArrays, ffi::PgvrxBuildStats)> {
        let st = self.index.build(&self.rows, &self.levels, Some(&self.heap_tids))?;
        self.rows = Vec::new();
        let g = self.index.export_graph()?;
        Ok((g, st))
    }

    /// The relation's pages for `g` (create_meta_page / create_graph_pages / write_neighbor_tuples / update_meta_page,
    /// build.rs:545-821): 8 KiB blocks, block 0 = meta page.
    pub fn write_pages(&self, g: &GraphArrays) -> Result<Vec<u8>> {
        // upper bound: every element on a page of its own plus the meta page
        let cap_pages = g.n + 2;
        let mut pages = vec![0u8; cap_pages as usize * 8192];
        let mut n_pages = 0i64;
        let rc = unsafe {
            ffi::pgvrx_pages_write(
                self.index.elem.code(),
                self.index.dim,
                self.index.m,
                self.index.ef_construction,
                g.n,
                g.rows.as_ptr() as *const c_void,
                g.level.as_ptr(),
                g.adj0.as_ptr(),
                g.upper_base.as_ptr(),
                g.adj_up.as_ptr(),
                g.n_upper_rows,
                g.entry,
                g.heap_tids.as_ptr(),
                g.heap_tid_counts.as_ptr(),
                pages.as_mut_ptr() as *mut c_void,
                cap_pages,
                &mut n_pages,
            )
        };
        if rc != ffi::PGVRX_OK {
            return Err(pgvrx_b200::Error { code: rc, message: "pgvrx_pages_write failed".into() });
        }
        pages.truncate(n_pages as usize * 8192);
        Ok(pages)
    }

    /// `aminsert` for a batch of rows into the finished (or loaded) index.
    pub fn insert_batch(&mut self, rows: &[u8], levels: &[i32], heap_tids: &[u64]) -> Result<ffi::PgvrxBuildStats> {
        self.index.insert(rows, levels, Some(heap_tids))
    }

    // ---- the narrow seam -------------------------------------------------------------------

    /// `search_layer` (graph/mod.rs:161-255) for a batch of (query, entry points) pairs on one layer: returns W sorted
    /// ascending per query plus the expanded-node log the caller's commit rule needs.
    #[allow(clippy::too_many_arguments)]
    pub fn search_layer_batch(
        &mut self,
        query_ids: &[i32],
        ep_ids: &[i32],
        ep_dists: &[f32],
        ep_counts: &[i32],
        ep_pitch: i32,
        ef: i32,
        layer: i32,
        exp_cap: i32,
    ) -> Result<(Vec<i32>, Vec<f32>, Vec<i32>, Vec<i32>, Vec<i32>)> {
        let nq = query_ids.len();
        let (mut ids, mut dists, mut counts) = (vec![-1i32; nq * ef as usize], vec![0f32; nq * ef as usize], vec![0i32; nq]);
        let (mut expanded, mut n_exp) = (vec![-1i32; nq * exp_cap as usize], vec![0i32; nq]);
        let rc = unsafe {
            ffi::pgvrx_search_layer_batch(
                self.index.raw(),
                std::ptr::null(),
                query_ids.as_ptr(),
                nq as i64,
                ep_ids.as_ptr(),
                ep_dists.as_ptr(),
                ep_counts.as_ptr(),
                ep_pitch,
                ef,
                layer,
                ids.as_mut_ptr(),
                dists.as_mut_ptr(),
                counts.as_mut_ptr(),
                expanded.as_mut_ptr(),
                exp_cap,
                n_exp.as_mut_ptr(),
            )
        };
        if rc != ffi::PGVRX_OK {
            return Err(pgvrx_b200::Error { code: rc, message: "pgvrx_search_layer_batch failed".into() });
        }
        Ok((ids, dists, counts, expanded, n_exp))
    }

    /// Keeps the device graph in step with a host-side `select_neighbors` / `update_neighbor_connections`: the new
    /// element's lists (layer 0 first with 2m slots, then layers 1..=level with m slots each, -1 after the last neighbour)
    /// and every neighbour list the host rewrote (list t belongs to (mod_nodes[t], mod_layers[t])).
    pub fn commit_insert(&mut self, row: &[u8], level: i32, lists: &[i32], mod_nodes: &[i32], mod_layers: &[i32], mod_lists: &[i32]) -> Result<i64> {
        let mut new_id = -1i64;
        let rc = unsafe {
            ffi::pgvrx_commit_insert(
                self.index.raw(),
                row.as_ptr() as *const c_void,
                level,
                lists.as_ptr(),
                mod_nodes.len() as i32,
                mod_nodes.as_ptr(),
                mod_layers.as_ptr(),
                mod_lists.as_ptr(),
                &mut new_id,
            )
        };
        if rc != ffi::PGVRX_OK {
            return Err(pgvrx_b200::Error { code: rc, message: "pgvrx_commit_insert failed".into() });
        }
        Ok(new_id)
    }
}
