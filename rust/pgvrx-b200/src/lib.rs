//! pgvrx-b200 -- safe wrapper over the C ABI in include/pgvrx_b200.h.
//!
//! This is the crate the pgvector-rx extension (rust/pgvector-rx-shim) and the standalone bench binary
//! (src/bin/bench.rs) call.  It owns nothing but handles: every buffer is the caller's, every failure is a `Result`
//! carrying the library's message (the library never unwinds; the pgrx side turns an `Err` into `pgrx::error!`, which
//! keeps PostgreSQL's longjmp on the Rust side of the boundary).
//!
//! No Rust toolchain exists in the image this repository is developed in: the crate ships as source, and the C ABI it
//! binds is what the repository's tests exercise (ctypes, C99, C++).  tests/test_rust_sys_cpu.py keeps the `-sys` crate in
//! step with the header and checks that every `ffi::` symbol used here exists there with the same arity.
use pgvrx_b200_sys as ffi;
use std::ffi::CStr;
use std::os::raw::c_void;
use std::ptr;

/// Status code + message of a failed call.
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}

impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "pgvrx error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}

pub type Result<T> = std::result::Result<T, Error>;

/// Element storage type (FUNCTION 3's `max_dimensions` identifies it in the reference, build.rs:133-175).
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Elem {
    F32,
    F16,
    Bit,
}

impl Elem {
    pub fn code(self) -> i32 {
        match self {
            Elem::F32 => ffi::PGVRX_ELEM_F32,
            Elem::F16 => ffi::PGVRX_ELEM_F16,
            Elem::Bit => ffi::PGVRX_ELEM_BIT,
        }
    }
}

/// Distance the opclass's FUNCTION 1 computes (the value the index orders by).
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Metric {
    L2Squared,
    NegInnerProduct,
    L1,
    Hamming,
    Jaccard,
}

impl Metric {
    pub fn code(self) -> i32 {
        match self {
            Metric::L2Squared => ffi::PGVRX_METRIC_L2SQ,
            Metric::NegInnerProduct => ffi::PGVRX_METRIC_NEG_IP,
            Metric::L1 => ffi::PGVRX_METRIC_L1,
            Metric::Hamming => ffi::PGVRX_METRIC_HAMMING,
            Metric::Jaccard => ffi::PGVRX_METRIC_JACCARD,
        }
    }
}

/// Operator class -> (element type, metric, normalise rows and queries).  Mirrors the opclass definitions
/// (vector.rs:820-865, halfvec.rs:1040-1073, bitvec.rs:193-237): cosine = negative inner product on unit rows.
pub fn opclass(name: &str) -> Option<(Elem, Metric, bool)> {
    Some(match name {
        "vector_l2_ops" => (Elem::F32, Metric::L2Squared, false),
        "vector_ip_ops" => (Elem::F32, Metric::NegInnerProduct, false),
        "vector_cosine_ops" => (Elem::F32, Metric::NegInnerProduct, true),
        "vector_l1_ops" => (Elem::F32, Metric::L1, false),
        "halfvec_l2_ops" => (Elem::F16, Metric::L2Squared, false),
        "halfvec_ip_ops" => (Elem::F16, Metric::NegInnerProduct, false),
        "halfvec_cosine_ops" => (Elem::F16, Metric::NegInnerProduct, true),
        "halfvec_l1_ops" => (Elem::F16, Metric::L1, false),
        "bit_hamming_ops" => (Elem::Bit, Metric::Hamming, false),
        "bit_jaccard_ops" => (Elem::Bit, Metric::Jaccard, false),
        _ => return None,
    })
}

/// hnsw.iterative_scan (options.rs:85,168-175)
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum IterativeScan {
    Off,
    RelaxedOrder,
    StrictOrder,
}

impl IterativeScan {
    pub fn code(self) -> i32 {
        match self {
            IterativeScan::Off => ffi::PGVRX_ITER_OFF,
            IterativeScan::RelaxedOrder => ffi::PGVRX_ITER_RELAXED,
            IterativeScan::StrictOrder => ffi::PGVRX_ITER_STRICT,
        }
    }
}

pub fn device_count() -> i32 {
    unsafe { ffi::pgvrx_device_count() }
}

/// Bytes of one element's raw data (f32[dim], u16[dim], ceil(dim / 8)).
pub fn row_bytes(elem: Elem, dim: i32) -> usize {
    unsafe { ffi::pgvrx_row_bytes(elem.code(), dim) }
}

/// Level of a new element from a uniform draw (build.rs:373-377, insert.rs:1298-1299); the caller owns the RNG.
pub fn level_from_uniform(r: f64, m: i32) -> i32 {
    unsafe { ffi::pgvrx_level_from_uniform(r, m) }
}

/// Cosine preparation of one row (vector.rs:106-126, halfvec.rs:204-233): writes the unit row, returns the f64 norm.
/// Rows whose norm is zero must not be indexed (build.rs:433-435).
pub fn normalize(elem: Elem, dim: i32, src: &[u8], out: &mut [u8]) -> Result<f64> {
    let rb = row_bytes(elem, dim);
    assert!(src.len() >= rb && out.len() >= rb);
    let mut norm = 0f64;
    let rc = unsafe { ffi::pgvrx_normalize(elem.code(), dim, src.as_ptr() as *const c_void, out.as_mut_ptr() as *mut c_void, &mut norm) };
    check_global(rc)?;
    Ok(norm)
}

fn message(idx: *const ffi::PgvrxIndex) -> String {
    let p = unsafe { ffi::pgvrx_last_error(idx) };
    if p.is_null() {
        String::new()
    } else {
        unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
    }
}

fn check_global(rc: i32) -> Result<()> {
    if rc == ffi::PGVRX_OK {
        Ok(())
    } else {
        Err(Error { code: rc, message: message(ptr::null()) })
    }
}

/// Graph + vector table as plain arrays: what `pgvrx_pages_export` produces from index pages and what
/// `Index::load_graph` takes.  Element id = page-walk order = insert order.
#[derive(Debug, Clone, Default)]
pub struct GraphArrays {
    pub n: i64,
    pub n_upper_rows: i64,
    pub entry: i64,
    pub rows: Vec<u8>,
    pub level: Vec<u8>,
    pub adj0: Vec<i32>,
    pub upper_base: Vec<i32>,
    pub adj_up: Vec<i32>,
    pub heap_tids: Vec<u64>,
    pub heap_tid_counts: Vec<u8>,
}

/// Walk a relation's pages (8 KiB blocks, block 0 = meta page) into arrays.  types/hnsw.rs:18-169 is the format.
pub fn export_pages(pages: &[u8], elem: Elem) -> Result<(ffi::PgvrxPagesInfo, GraphArrays, Vec<u8>)> {
    let n_pages = (pages.len() / 8192) as i64;
    let mut info: ffi::PgvrxPagesInfo = unsafe { std::mem::zeroed() };
    check_global(unsafe { ffi::pgvrx_pages_inspect(pages.as_ptr() as *const c_void, n_pages, elem.code(), &mut info) })?;
    let n = info.n_elements as usize;
    let rb = row_bytes(elem, info.dims);
    let m = info.m as usize;
    let mut g = GraphArrays {
        n: info.n_elements,
        n_upper_rows: info.n_upper_rows,
        entry: info.entry,
        rows: vec![0u8; n * rb],
        level: vec![0u8; n],
        adj0: vec![-1i32; n * 2 * m],
        upper_base: vec![-1i32; n],
        adj_up: vec![-1i32; info.n_upper_rows as usize * m],
        heap_tids: vec![!0u64; n * ffi::PGVRX_HEAPTIDS as usize],
        heap_tid_counts: vec![0u8; n],
    };
    check_global(unsafe {
        ffi::pgvrx_pages_export(
            pages.as_ptr() as *const c_void,
            n_pages,
            elem.code(),
            g.rows.as_mut_ptr() as *mut c_void,
            g.level.as_mut_ptr(),
            g.adj0.as_mut_ptr(),
            g.upper_base.as_mut_ptr(),
            g.adj_up.as_mut_ptr(),
            g.heap_tids.as_mut_ptr(),
            g.heap_tid_counts.as_mut_ptr(),
        )
    })?;
    let mut flags = vec![0u8; n];
    check_global(unsafe { ffi::pgvrx_pages_element_flags(pages.as_ptr() as *const c_void, n_pages, flags.as_mut_ptr()) })?;
    Ok((info, g, flags))
}

/// Heap TIDs in the order `amgettuple` hands them out for a list of element ids: the last TID of every element first
/// (scan.rs:798).
pub fn heap_tid_stream(elem_ids: &[i32], g: &GraphArrays) -> Result<Vec<u64>> {
    let mut out = vec![0u64; elem_ids.len() * ffi::PGVRX_HEAPTIDS as usize];
    let mut n_out = 0i64;
    check_global(unsafe {
        ffi::pgvrx_heap_tid_stream(
            elem_ids.as_ptr(),
            elem_ids.len() as i64,
            g.heap_tids.as_ptr(),
            g.heap_tid_counts.as_ptr(),
            g.n,
            out.as_mut_ptr(),
            out.len() as i64,
            &mut n_out,
        )
    })?;
    out.truncate(n_out as usize);
    Ok(out)
}

/// One index resident on one GPU.
pub struct Index {
    raw: *mut ffi::PgvrxIndex,
    pub elem: Elem,
    pub metric: Metric,
    pub dim: i32,
    pub m: i32,
    pub ef_construction: i32,
}

// One PostgreSQL backend = one thread (build.rs:380-385); a handle is used by one thread at a time but may move.
unsafe impl Send for Index {}

/// Result rows of a batch: `k` slots per query, padded with -1 / +inf; `counts[q]` valid entries, nearest first.
#[derive(Debug, Clone)]
pub struct BatchResult {
    pub k: usize,
    pub ids: Vec<i32>,
    pub dists: Vec<f32>,
    pub counts: Vec<i32>,
}

impl BatchResult {
    pub fn row(&self, q: usize) -> (&[i32], &[f32]) {
        let c = self.counts[q].max(0) as usize;
        (&self.ids[q * self.k..q * self.k + c], &self.dists[q * self.k..q * self.k + c])
    }
}

impl Index {
    /// `capacity` = the most elements the handle will ever hold (device arrays are sized once).  Option bounds are the
    /// reloption bounds (options.rs:203-225) and fail with the reference's own ERROR text.
    pub fn create(elem: Elem, metric: Metric, dim: i32, m: i32, ef_construction: i32, device: i32, capacity: i64) -> Result<Index> {
        let desc = ffi::PgvrxIndexDesc { elem_type: elem.code(), metric: metric.code(), dim, m, ef_construction, device, capacity };
        let mut raw: *mut ffi::PgvrxIndex = ptr::null_mut();
        let rc = unsafe { ffi::pgvrx_index_create(&desc, &mut raw) };
        check_global(rc)?;
        Ok(Index { raw, elem, metric, dim, m, ef_construction })
    }

    pub fn raw(&self) -> *mut ffi::PgvrxIndex {
        self.raw
    }

    fn check(&self, rc: i32) -> Result<()> {
        if rc == ffi::PGVRX_OK {
            Ok(())
        } else {
            Err(Error { code: rc, message: message(self.raw) })
        }
    }

    pub fn row_bytes(&self) -> usize {
        row_bytes(self.elem, self.dim)
    }

    pub fn len(&self) -> i64 {
        unsafe { ffi::pgvrx_index_size(self.raw) }
    }

    /// Graph exported from pages (or from another handle) into HBM.
    pub fn load_graph(&mut self, g: &GraphArrays) -> Result<()> {
        self.check(unsafe {
            ffi::pgvrx_index_load_graph(
                self.raw,
                g.rows.as_ptr() as *const c_void,
                g.n,
                g.level.as_ptr(),
                g.adj0.as_ptr(),
                g.upper_base.as_ptr(),
                g.adj_up.as_ptr(),
                g.n_upper_rows,
                g.entry,
            )
        })?;
        if !g.heap_tid_counts.is_empty() {
            self.check(unsafe { ffi::pgvrx_index_set_heap_tids(self.raw, g.heap_tids.as_ptr(), g.heap_tid_counts.as_ptr(), g.n) })?;
        }
        Ok(())
    }

    /// Per-element vacuum state of a relation that `ambulkdelete` has worked on (scan.rs:179, 866-868).
    pub fn set_element_flags(&mut self, flags: &[u8]) -> Result<()> {
        self.check(unsafe { ffi::pgvrx_index_set_element_flags(self.raw, flags.as_ptr(), flags.len() as i64) })
    }

    /// `ambuild`'s insert loop (build.rs:400-535) for rows in heap-scan order; `levels[i]` from `level_from_uniform`.
    pub fn build(&mut self, rows: &[u8], levels: &[i32], heap_tids: Option<&[u64]>) -> Result<ffi::PgvrxBuildStats> {
        let n = levels.len() as i64;
        assert!(rows.len() >= levels.len() * self.row_bytes());
        let mut st: ffi::PgvrxBuildStats = unsafe { std::mem::zeroed() };
        self.check(unsafe {
            ffi::pgvrx_index_build(self.raw, rows.as_ptr() as *const c_void, n, levels.as_ptr(), heap_tids.map_or(ptr::null(), |t| t.as_ptr()), 0, &mut st)
        })?;
        Ok(st)
    }

    /// `aminsert`'s rules (insert.rs:1021-1123, 500-738) for a batch of rows, equal to inserting them one at a time.
    pub fn insert(&mut self, rows: &[u8], levels: &[i32], heap_tids: Option<&[u64]>) -> Result<ffi::PgvrxBuildStats> {
        let n = levels.len() as i64;
        assert!(rows.len() >= levels.len() * self.row_bytes());
        let mut st: ffi::PgvrxBuildStats = unsafe { std::mem::zeroed() };
        self.check(unsafe {
            ffi::pgvrx_index_insert(self.raw, rows.as_ptr() as *const c_void, n, levels.as_ptr(), heap_tids.map_or(ptr::null(), |t| t.as_ptr()), 0, &mut st)
        })?;
        Ok(st)
    }

    /// `ambulkdelete`'s three passes (vacuum.rs:816-842) for a set of dead heap TIDs.
    pub fn vacuum(&mut self, dead_tids: &[u64]) -> Result<ffi::PgvrxVacuumStats> {
        let mut st: ffi::PgvrxVacuumStats = unsafe { std::mem::zeroed() };
        self.check(unsafe { ffi::pgvrx_index_vacuum(self.raw, dead_tids.as_ptr(), dead_tids.len() as i64, 0, &mut st) })?;
        Ok(st)
    }

    /// The graph as it stands (after build / insert / vacuum): what the flush phase (build.rs:545-821) writes to pages.
    pub fn export_graph(&mut self) -> Result<GraphArrays> {
        let (mut n, mut nu, mut entry, mut el) = (0i64, 0i64, 0i64, 0i32);
        self.check(unsafe { ffi::pgvrx_index_graph_info(self.raw, &mut n, &mut nu, &mut entry, &mut el) })?;
        let (nn, m) = (n as usize, self.m as usize);
        let mut g = GraphArrays {
            n,
            n_upper_rows: nu,
            entry,
            rows: vec![0u8; nn * self.row_bytes()],
            level: vec![0u8; nn],
            adj0: vec![-1i32; nn * 2 * m],
            upper_base: vec![-1i32; nn],
            adj_up: vec![-1i32; nu as usize * m],
            heap_tids: vec![!0u64; nn * ffi::PGVRX_HEAPTIDS as usize],
            heap_tid_counts: vec![0u8; nn],
        };
        self.check(unsafe {
            ffi::pgvrx_index_export_graph(
                self.raw,
                g.rows.as_mut_ptr() as *mut c_void,
                g.level.as_mut_ptr(),
                g.adj0.as_mut_ptr(),
                g.upper_base.as_mut_ptr(),
                g.adj_up.as_mut_ptr(),
                g.heap_tids.as_mut_ptr(),
                g.heap_tid_counts.as_mut_ptr(),
            )
        })?;
        Ok(g)
    }

    /// `get_scan_items` (scan.rs:458-530) for `nq` queries (raw element data, normalised already for cosine).
    pub fn search(&mut self, queries: &[u8], nq: usize, ef_search: i32, k: i32) -> Result<(BatchResult, ffi::PgvrxSearchStats)> {
        assert!(queries.len() >= nq * self.row_bytes());
        let ku = k as usize;
        let mut r = BatchResult { k: ku, ids: vec![-1; nq * ku], dists: vec![f32::INFINITY; nq * ku], counts: vec![0; nq] };
        let mut st: ffi::PgvrxSearchStats = unsafe { std::mem::zeroed() };
        self.check(unsafe {
            ffi::pgvrx_search_batch(
                self.raw,
                queries.as_ptr() as *const c_void,
                nq as i64,
                ef_search,
                k,
                r.ids.as_mut_ptr(),
                r.dists.as_mut_ptr(),
                r.counts.as_mut_ptr(),
                &mut st,
            )
        })?;
        Ok((r, st))
    }

    /// Operator-ordered sequential scan (the reference's recall ground truth) on the GPU: exact, ordered by (distance, id).
    pub fn flat_topk(&mut self, queries: &[u8], nq: usize, k: i32) -> Result<(Vec<i32>, Vec<f32>, ffi::PgvrxFlatStats)> {
        let ku = k as usize;
        let (mut ids, mut dists) = (vec![-1i32; nq * ku], vec![f32::INFINITY; nq * ku]);
        let mut st: ffi::PgvrxFlatStats = unsafe { std::mem::zeroed() };
        self.check(unsafe { ffi::pgvrx_flat_topk(self.raw, queries.as_ptr() as *const c_void, nq as i64, k, ids.as_mut_ptr(), dists.as_mut_ptr(), &mut st) })?;
        Ok((ids, dists, st))
    }

    /// Exact re-rank of caller-supplied candidate lists (`nc` ids per query, -1 = none).
    pub fn rerank(&mut self, queries: &[u8], nq: usize, cand: &[i32], nc: i32, k: i32) -> Result<(Vec<i32>, Vec<f32>)> {
        let ku = k as usize;
        let (mut ids, mut dists) = (vec![-1i32; nq * ku], vec![f32::INFINITY; nq * ku]);
        self.check(unsafe { ffi::pgvrx_rerank(self.raw, queries.as_ptr() as *const c_void, nq as i64, cand.as_ptr(), nc, k, ids.as_mut_ptr(), dists.as_mut_ptr()) })?;
        Ok((ids, dists))
    }

    /// Iterative scan state for a batch of queries (HnswScanState's visited / discarded / tuples, scan.rs:584-613, in HBM).
    pub fn scan_begin(&mut self, queries: &[u8], nq: usize, ef_search: i32, mode: IterativeScan, max_scan_tuples: i64) -> Result<Scan<'_>> {
        let mut raw: *mut ffi::PgvrxScan = ptr::null_mut();
        self.check(unsafe { ffi::pgvrx_scan_begin(self.raw, queries.as_ptr() as *const c_void, nq as i64, ef_search, mode.code(), max_scan_tuples, &mut raw) })?;
        Ok(Scan { raw, index: self, nq, ef: ef_search as usize })
    }
}

impl Drop for Index {
    fn drop(&mut self) {
        unsafe { ffi::pgvrx_index_destroy(self.raw) }
    }
}

/// An open iterative scan.  Every `next` is one refill of `so.results` (scan.rs:817-856) for all queries still wanted.
pub struct Scan<'a> {
    raw: *mut ffi::PgvrxScan,
    index: &'a mut Index,
    nq: usize,
    ef: usize,
}

impl<'a> Scan<'a> {
    /// `want_more[q] == 0` skips query q this round.  Returns up to `ef_search` element ids per query in emission order
    /// (strict-order filter and the `max_scan_tuples` drain applied) and, per query, whether the scan is exhausted.
    pub fn next(&mut self, want_more: Option<&[u8]>) -> Result<(BatchResult, Vec<u8>)> {
        let mut r = BatchResult { k: self.ef, ids: vec![-1; self.nq * self.ef], dists: vec![f32::INFINITY; self.nq * self.ef], counts: vec![0; self.nq] };
        let mut done = vec![0u8; self.nq];
        let rc = unsafe {
            ffi::pgvrx_scan_next(
                self.raw,
                want_more.map_or(ptr::null(), |w| w.as_ptr()),
                r.ids.as_mut_ptr(),
                r.dists.as_mut_ptr(),
                r.counts.as_mut_ptr(),
                done.as_mut_ptr(),
                ptr::null_mut(),
            )
        };
        self.index.check(rc)?;
        Ok((r, done))
    }
}

impl<'a> Scan<'a> {
    /// `rounds` refills in one launch (batch r of query q at `(q * rounds + r) * ef`): the same stream as `rounds` calls
    /// of `next`, without the per-call launch and copy -- for consumers under a selective filter.
    pub fn next_rounds(&mut self, want_more: Option<&[u8]>, rounds: i32) -> Result<(Vec<i32>, Vec<f32>, Vec<i32>, Vec<u8>)> {
        let n = self.nq * rounds as usize;
        let (mut ids, mut dists, mut counts) = (vec![-1i32; n * self.ef], vec![f32::INFINITY; n * self.ef], vec![0i32; n]);
        let mut done = vec![0u8; self.nq];
        let rc = unsafe {
            ffi::pgvrx_scan_next_rounds(
                self.raw,
                want_more.map_or(ptr::null(), |w| w.as_ptr()),
                rounds,
                ids.as_mut_ptr(),
                dists.as_mut_ptr(),
                counts.as_mut_ptr(),
                done.as_mut_ptr(),
                ptr::null_mut(),
            )
        };
        self.index.check(rc)?;
        Ok((ids, dists, counts, done))
    }
}

impl<'a> Drop for Scan<'a> {
    fn drop(&mut self) {
        unsafe { ffi::pgvrx_scan_end(self.raw) }
    }
}

/// Several GPUs of one box behind one handle: query-sharded scan over replicas, row-sharded exact kNN with the merge
/// reading the peers' lists over NVLink.
pub struct Group {
    raw: *mut ffi::PgvrxGroup,
    row_bytes: usize,
}

unsafe impl Send for Group {}

impl Group {
    pub fn create(elem: Elem, metric: Metric, dim: i32, m: i32, ef_construction: i32, devices: &[i32], capacity: i64) -> Result<Group> {
        let desc = ffi::PgvrxIndexDesc { elem_type: elem.code(), metric: metric.code(), dim, m, ef_construction, device: 0, capacity };
        let mut raw: *mut ffi::PgvrxGroup = ptr::null_mut();
        let rc = unsafe { ffi::pgvrx_group_create(&desc, devices.as_ptr(), devices.len() as i32, &mut raw) };
        if rc != ffi::PGVRX_OK {
            let p = unsafe { ffi::pgvrx_group_last_error(ptr::null()) };
            let message = if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() };
            return Err(Error { code: rc, message });
        }
        Ok(Group { raw, row_bytes: row_bytes(elem, dim) })
    }

    fn check(&self, rc: i32) -> Result<()> {
        if rc == ffi::PGVRX_OK {
            return Ok(());
        }
        let p = unsafe { ffi::pgvrx_group_last_error(self.raw) };
        let message = if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() };
        Err(Error { code: rc, message })
    }

    pub fn size(&self) -> i32 {
        unsafe { ffi::pgvrx_group_size(self.raw) }
    }

    pub fn load_graph(&mut self, g: &GraphArrays) -> Result<()> {
        self.check(unsafe {
            ffi::pgvrx_group_load_graph(
                self.raw,
                g.rows.as_ptr() as *const c_void,
                g.n,
                g.level.as_ptr(),
                g.adj0.as_ptr(),
                g.upper_base.as_ptr(),
                g.adj_up.as_ptr(),
                g.n_upper_rows,
                g.entry,
            )
        })
    }

    /// One caller batch split over the devices; every device writes its slice of the result arrays.
    pub fn search(&mut self, queries: &[u8], nq: usize, ef_search: i32, k: i32) -> Result<(BatchResult, ffi::PgvrxGroupStats)> {
        assert!(queries.len() >= nq * self.row_bytes);
        let ku = k as usize;
        let mut r = BatchResult { k: ku, ids: vec![-1; nq * ku], dists: vec![f32::INFINITY; nq * ku], counts: vec![0; nq] };
        let mut st: ffi::PgvrxGroupStats = unsafe { std::mem::zeroed() };
        self.check(unsafe {
            ffi::pgvrx_group_search_batch(self.raw, queries.as_ptr() as *const c_void, nq as i64, ef_search, k, r.ids.as_mut_ptr(), r.dists.as_mut_ptr(), r.counts.as_mut_ptr(), &mut st)
        })?;
        Ok((r, st))
    }

    pub fn load_rows(&mut self, rows: &[u8], n: i64) -> Result<()> {
        self.check(unsafe { ffi::pgvrx_group_load_rows(self.raw, rows.as_ptr() as *const c_void, n) })
    }

    pub fn flat_topk(&mut self, queries: &[u8], nq: usize, k: i32) -> Result<(Vec<i32>, Vec<f32>, ffi::PgvrxGroupStats)> {
        let ku = k as usize;
        let (mut ids, mut dists) = (vec![-1i32; nq * ku], vec![f32::INFINITY; nq * ku]);
        let mut st: ffi::PgvrxGroupStats = unsafe { std::mem::zeroed() };
        self.check(unsafe { ffi::pgvrx_group_flat_topk(self.raw, queries.as_ptr() as *const c_void, nq as i64, k, ids.as_mut_ptr(), dists.as_mut_ptr(), &mut st) })?;
        Ok((ids, dists, st))
    }
}

impl Drop for Group {
    fn drop(&mut self) {
        unsafe { ffi::pgvrx_group_destroy(self.raw) }
    }
}
