//! GUCs and reloptions as the bridges consume them.  Registration stays in src/index/options.rs (`init_gucs`,
//! `init_relopts`, options.rs:156-258); these are the values read there, with the same names, bounds and defaults, so a
//! maintainer passes `HNSW_EF_SEARCH.get()` etc. straight through.
use pgvrx_b200::IterativeScan;

/// hnsw.ef_search 1..1000, default 40 (options.rs:81,157-166)
pub const EF_SEARCH_DEFAULT: i32 = 40;
pub const EF_SEARCH_MIN: i32 = 1;
pub const EF_SEARCH_MAX: i32 = 1000;
/// hnsw.max_scan_tuples, default 20000 (options.rs:90)
pub const MAX_SCAN_TUPLES_DEFAULT: i64 = 20_000;
/// reloption m 2..100, default 16; ef_construction 4..1000, default 64 (options.rs:203-225)
pub const M_DEFAULT: i32 = 16;
pub const EF_CONSTRUCTION_DEFAULT: i32 = 64;

#[derive(Debug, Clone, Copy)]
pub struct ScanGucs {
    pub ef_search: i32,
    pub iterative_scan: IterativeScan,
    pub max_scan_tuples: i64,
}

impl Default for ScanGucs {
    fn default() -> Self {
        ScanGucs { ef_search: EF_SEARCH_DEFAULT, iterative_scan: IterativeScan::Off, max_scan_tuples: MAX_SCAN_TUPLES_DEFAULT }
    }
}

/// `hnsw.iterative_scan` enum value as options.rs:168-175 orders it (off, relaxed_order, strict_order).
pub fn iterative_from_guc(v: i32) -> IterativeScan {
    match v {
        1 => IterativeScan::RelaxedOrder,
        2 => IterativeScan::StrictOrder,
        _ => IterativeScan::Off,
    }
}
