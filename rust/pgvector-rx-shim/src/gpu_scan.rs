//! Scan seam.
//!
//! Call-site change in src/index/scan.rs (`amgettuple`, first call, scan.rs:717-791): where the reference calls
//! `get_scan_items(index, query, dist_fmgr, collation, ef, visited, discarded)` and stores the candidates in
//! `so.results`, call `ScanBridge::first(...)` and store what it returns; where it refills `so.results` through
//! `resume_scan_items` (scan.rs:817-856), call `ScanBridge::refill(...)`.  Everything after that is unchanged: the
//! reference pops `so.results` from the back (nearest last), hands out each element's heap TIDs last-to-first
//! (scan.rs:794-875) and sets `xs_recheck = false`.
//!
//! The bridge keeps, per index, the exporter's element-id -> (index TID, heap TIDs) table, so the GPU's element ids
//! turn back into the `ScanCandidate`s the rest of `amgettuple` expects.
use crate::guc::ScanGucs;
use pgvrx_b200::{export_pages, Elem, Error, GraphArrays, Index, IterativeScan, Metric, Result};

/// What `amgettuple` needs per result element: distance (f64, as ScanCandidate.distance), the element's heap TIDs in
/// stored order (the caller pops from the back, scan.rs:798).
#[derive(Debug, Clone)]
pub struct GpuCandidate {
    pub element_id: i32,
    pub distance: f64,
    pub heaptids: Vec<u64>,
}

/// One index relation resident on the GPU.  Lives in a per-backend cache keyed by the index OID (or in a sidecar that
/// owns the CUDA context for all backends -- the ABI is the same, `nq` is just larger).
pub struct ScanBridge {
    index: Index,
    graph_tids: GraphArrays,          // only heap_tids / heap_tid_counts are kept (rows and lists live in HBM)
}

impl ScanBridge {
    /// `pages`: the relation's blocks in block order (ReadBuffer + memcpy of every block, or the relation file itself).
    /// Re-run after VACUUM or after inserts that bypassed the bridge: the GPU copy is a snapshot of the relation.
    pub fn from_pages(pages: &[u8], elem: Elem, metric: Metric, device: i32) -> Result<ScanBridge> {
        let (info, mut g, flags) = export_pages(pages, elem)?;
        let mut index = Index::create(elem, metric, info.dims, info.m, info.ef_construction, device, info.n_elements.max(1))?;
        index.load_graph(&g)?;
        if flags.iter().any(|&f| f != 0) {
            index.set_element_flags(&flags)?;        // deleted / no-heap-TID elements (scan.rs:179, 866-868)
        }
        g.rows = Vec::new();
        g.adj0 = Vec::new();
        g.adj_up = Vec::new();
        Ok(ScanBridge { index, graph_tids: g })
    }

    fn candidates(&self, ids: &[i32], dists: &[f32]) -> Vec<GpuCandidate> {
        // nearest LAST, as get_scan_items returns them (scan.rs:441-447): amgettuple pops from the back
        ids.iter()
            .zip(dists)
            .rev()
            .map(|(&id, &d)| {
                let c = self.graph_tids.heap_tid_counts[id as usize] as usize;
                let base = id as usize * 10;
                GpuCandidate { element_id: id, distance: d as f64, heaptids: self.graph_tids.heap_tids[base..base + c].to_vec() }
            })
            .collect()
    }

    /// Replaces `get_scan_items` for a non-iterative scan: `query` is the varlena payload, normalised already for cosine
    /// opclasses (scan.rs:749-754).  At most `ef_search` elements, as in the reference.
    pub fn first(&mut self, query: &[u8], gucs: &ScanGucs) -> Result<Vec<GpuCandidate>> {
        let (r, _) = self.index.search(query, 1, gucs.ef_search, gucs.ef_search)?;
        let (ids, dists) = r.row(0);
        Ok(self.candidates(ids, dists))
    }

    /// Iterative scans: the state the reference keeps in `HnswScanState{visited, discarded, tuples, previous_distance}`
    /// (scan.rs:584-613) lives behind the returned scan; call `IterState::refill` wherever `so.results` runs empty.
    pub fn begin_iterative<'a>(&'a mut self, query: &[u8], gucs: &ScanGucs) -> Result<IterState<'a>> {
        if gucs.iterative_scan == IterativeScan::Off {
            return Err(Error { code: 1, message: "begin_iterative needs hnsw.iterative_scan != off".into() });
        }
        let tids = &self.graph_tids as *const GraphArrays;
        let scan = self.index.scan_begin(query, 1, gucs.ef_search, gucs.iterative_scan, gucs.max_scan_tuples)?;
        Ok(IterState { scan, tids })
    }
}

pub struct IterState<'a> {
    scan: pgvrx_b200::Scan<'a>,
    tids: *const GraphArrays,
}

impl<'a> IterState<'a> {
    /// Next batch of elements in emission order (strict-order filter and the max_scan_tuples drain already applied),
    /// nearest LAST; `true` when the scan is exhausted.
    pub fn refill(&mut self) -> Result<(Vec<GpuCandidate>, bool)> {
        let (r, done) = self.scan.next(None)?;
        let (ids, dists) = r.row(0);
        let g = unsafe { &*self.tids };
        let out = ids
            .iter()
            .zip(dists)
            .rev()
            .map(|(&id, &d)| {
                let c = g.heap_tid_counts[id as usize] as usize;
                let base = id as usize * 10;
                GpuCandidate { element_id: id, distance: d as f64, heaptids: g.heap_tids[base..base + c].to_vec() }
            })
            .collect();
        Ok((out, done[0] != 0))
    }
}
