"""ctypes binding of oracle/liborc.so -- the CPU restatement of the reference's HNSW path.

TEST INFRASTRUCTURE.  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never touches it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liborc.so")

ELEM_F32, ELEM_F16, ELEM_BIT = 0, 1, 2
L2SQ, NEG_IP, L1, HAMMING, JACCARD = 0, 1, 2, 3, 4
TIE_RUST_HEAP, TIE_TOTAL_ORDER = 0, 1
ITER_OFF, ITER_RELAXED, ITER_STRICT = 0, 1, 2
VACUUM_REMOVE_HEAPTIDS, VACUUM_MARK_DELETED, VACUUM_STALE_NEIGHBORS = 1, 2, 3
ELEM_DELETED, ELEM_NO_HEAPTIDS = 1, 2


class Counters(C.Structure):
    _fields_ = [("n_dist", C.c_uint64), ("n_expand0", C.c_uint64), ("n_expand_upper", C.c_uint64),
                ("n_visited", C.c_uint64), ("n_pair", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def build_oracle(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("hnsw_oracle.cpp", "hnsw_oracle.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


_lib = None


def build_native():
    """liborc_native.so: the same source, `-O3 -march=native`, compiled on the machine that runs this (bench.py's CPU
    legs set PGVRX_ORACLE_NATIVE=1: BASELINE.md promises these flags for the timed baseline).  Falls back to the portable
    build when there is no compiler."""
    out = os.path.join(ORACLE_DIR, "liborc_native.so")
    try:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "-B", "native"])
        return out
    except (OSError, subprocess.CalledProcessError):
        return None


NATIVE = False        # True once lib() has loaded the -march=native build


def lib():
    global _lib, NATIVE
    if _lib is not None:
        return _lib
    build_oracle()
    path = LIB_PATH
    if os.environ.get("PGVRX_ORACLE_NATIVE") == "1":
        p = build_native()
        if p:
            path, NATIVE = p, True
    L = C.CDLL(path)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    sig = {
        "orc_half_to_f32": (f32, [C.c_uint16]),
        "orc_f32_to_half": (C.c_uint16, [f32]),
        "orc_distance": (f64, [i32, i32, i32, vp, vp]),
        "orc_normalize": (f64, [i32, i32, vp, vp]),
        "orc_level_from_uniform": (i32, [f64, i32]),
        "orc_max_level": (i32, [i32]),
        "orc_create": (vp, [i32] * 6),
        "orc_destroy": (None, [vp]),
        "orc_row_bytes": (C.c_size_t, [vp]),
        "orc_build": (i64, [vp, vp, i64, vp, i64]),
        "orc_insert": (i64, [vp, vp, i64, vp, i64]),
        "orc_n_elements": (i64, [vp]),
        "orc_entry_point": (i64, [vp]),
        "orc_element_level": (i32, [vp, i64]),
        "orc_neighbors": (i32, [vp, i64, i32, vp, vp, i32]),
        "orc_heap_tids": (i32, [vp, i64, vp, i32]),
        "orc_element_row": (vp, [vp, i64]),
        "orc_build_counters": (None, [vp, vp]),
        "orc_export_graph": (i64, [vp, vp, vp, vp, vp]),
        "orc_import_graph": (i32, [vp, vp, i64, vp, vp, vp, vp, i64]),
        "orc_search_layer": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp]),
        "orc_select_neighbors": (i32, [vp, vp, vp, i32, i32, vp, vp]),
        "orc_search_batch": (i32, [vp, vp, i64, i32, i32, vp, vp, vp, vp, i32]),
        "orc_scan_stream": (i64, [vp, vp, i32, i32, i64, i64, vp, vp, vp, vp]),
        "orc_flat_topk": (i32, [i32, i32, i32, vp, i64, vp, i64, i32, vp, vp, i32]),
        "orc_write_pages": (i64, [vp, vp, i64]),
        "orc_vacuum_mark": (i32, [vp, i64, i32]),
        "orc_bulk_delete": (i64, [vp, vp, i64]),
        "orc_element_flags": (None, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def elem_dtype(elem_type):
    return {ELEM_F32: np.float32, ELEM_F16: np.uint16, ELEM_BIT: np.uint8}[elem_type]


def row_elems(elem_type, dim):
    return (dim + 7) // 8 if elem_type == ELEM_BIT else dim


def as_rows(elem_type, dim, a):
    """Return a C-contiguous 2-D array of raw element data (f32 / u16 half bits / packed bytes)."""
    a = np.asarray(a)
    if elem_type == ELEM_F16 and a.dtype == np.float16:
        a = a.view(np.uint16)
    a = np.ascontiguousarray(a, dtype=elem_dtype(elem_type))
    if a.ndim == 1:
        a = a[None, :]
    assert a.shape[1] == row_elems(elem_type, dim), (a.shape, dim)
    return a


def distance(elem_type, metric, dim, a, b):
    a = as_rows(elem_type, dim, a)
    b = as_rows(elem_type, dim, b)
    return lib().orc_distance(elem_type, metric, dim, _p(a), _p(b))


def normalize(elem_type, dim, a):
    a = as_rows(elem_type, dim, a)
    out = np.zeros_like(a)
    norm = lib().orc_normalize(elem_type, dim, _p(a), _p(out))
    return norm, out[0]


def levels_from_uniform(r, m):
    L = lib()
    return np.array([L.orc_level_from_uniform(float(x), m) for x in r], dtype=np.int32)


class Oracle:
    def __init__(self, elem_type, metric, dim, m=16, ef_construction=64, tie_mode=TIE_TOTAL_ORDER):
        self.L = lib()
        self.elem_type, self.metric, self.dim, self.m, self.efc = elem_type, metric, dim, m, ef_construction
        self.h = self.L.orc_create(elem_type, metric, dim, m, ef_construction, tie_mode)
        if not self.h:
            raise ValueError("invalid index options")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    # ---- build -------------------------------------------------------------------------
    def build(self, rows, levels, base_tid=0):
        rows = as_rows(self.elem_type, self.dim, rows)
        levels = np.ascontiguousarray(levels, dtype=np.int32)
        assert len(levels) == len(rows)
        return self.L.orc_build(self.h, _p(rows), len(rows), _p(levels), base_tid)

    def insert(self, rows, levels, base_tid=0):
        """aminsert for every row in order (src/index/insert.rs:1227-1480)."""
        rows = as_rows(self.elem_type, self.dim, rows)
        levels = np.ascontiguousarray(levels, dtype=np.int32)
        assert len(levels) == len(rows)
        return self.L.orc_insert(self.h, _p(rows), len(rows), _p(levels), base_tid)

    @property
    def n(self):
        return self.L.orc_n_elements(self.h)

    @property
    def entry(self):
        return self.L.orc_entry_point(self.h)

    def level(self, i):
        return self.L.orc_element_level(self.h, i)

    def neighbors(self, i, layer, with_dists=False):
        cap = 2 * self.m
        ids = np.empty(cap, np.int32)
        ds = np.empty(cap, np.float32)
        k = self.L.orc_neighbors(self.h, i, layer, _p(ids), _p(ds), cap)
        return (ids[:k].copy(), ds[:k].copy()) if with_dists else ids[:k].copy()

    def heap_tids(self, i):
        t = np.empty(10, np.int64)
        k = self.L.orc_heap_tids(self.h, i, _p(t), 10)
        return t[:k].copy()

    def build_counters(self):
        c = Counters()
        self.L.orc_build_counters(self.h, C.byref(c))
        return c.as_dict()

    def export_graph(self):
        n, m = self.n, self.m
        n_up = self.L.orc_export_graph(self.h, None, None, None, None)
        level = np.empty(n, np.uint8)
        adj0 = np.empty((n, 2 * m), np.int32)
        upper_base = np.empty(n, np.int32)
        adj_up = np.empty((max(n_up, 1), m), np.int32)
        self.L.orc_export_graph(self.h, _p(level), _p(adj0), _p(upper_base), _p(adj_up))
        rb = self.L.orc_row_bytes(self.h)
        rows = np.empty((n, rb), np.uint8)
        if n:
            base = self.L.orc_element_row(self.h, 0)
            C.memmove(rows.ctypes.data, base, n * rb)
        rows = rows.view(elem_dtype(self.elem_type)).reshape(n, -1)
        return dict(rows=rows, level=level, adj0=adj0, upper_base=upper_base, adj_up=adj_up[:n_up],
                    entry=self.entry, n=n)

    def import_graph(self, g):
        rows = as_rows(self.elem_type, self.dim, g["rows"])
        adj_up = np.ascontiguousarray(g["adj_up"], np.int32)
        if adj_up.size == 0:
            adj_up = np.zeros((1, self.m), np.int32)
        self._keep = (rows, adj_up)
        self.L.orc_import_graph(self.h, _p(rows), len(rows), _p(np.ascontiguousarray(g["level"], np.uint8)),
                                _p(np.ascontiguousarray(g["adj0"], np.int32)),
                                _p(np.ascontiguousarray(g["upper_base"], np.int32)), _p(adj_up),
                                int(g["entry"]))

    def bulk_delete(self, dead_tids):
        """VACUUM: ambulkdelete with the given dead heap TIDs (src/index/vacuum.rs:816-842)."""
        t = np.ascontiguousarray(dead_tids, np.int64)
        return self.L.orc_bulk_delete(self.h, _p(t), len(t))

    def vacuum_mark(self, ids, what):
        """Put elements into a state a vacuum leaves on disk (VACUUM_* constants)."""
        for i in np.atleast_1d(ids):
            assert self.L.orc_vacuum_mark(self.h, int(i), what) == 0

    def element_flags(self):
        f = np.zeros(self.n, np.uint8)
        self.L.orc_element_flags(self.h, _p(f))
        return f

    def write_pages(self):
        """The index relation as the reference's flush phase lays it out (8 KiB pages)."""
        n = self.L.orc_write_pages(self.h, None, 0)
        buf = np.zeros(n * 8192, np.uint8)
        assert self.L.orc_write_pages(self.h, _p(buf), n) == n
        return buf

    # ---- a1 / a3 -----------------------------------------------------------------------
    def search_layer(self, query, ep_ids, ep_dists, ef, layer):
        q = as_rows(self.elem_type, self.dim, query)
        ep_ids = np.ascontiguousarray(ep_ids, np.int32)
        ep_dists = np.ascontiguousarray(ep_dists, np.float32)
        cap = ef + len(ep_ids) + 1
        ids = np.empty(cap, np.int32)
        ds = np.empty(cap, np.float32)
        k = self.L.orc_search_layer(self.h, _p(q), _p(ep_ids), _p(ep_dists), len(ep_ids), ef, layer,
                                    _p(ids), _p(ds))
        return ids[:k].copy(), ds[:k].copy()

    def select_neighbors(self, cand_ids, cand_dists, max_neighbors):
        cand_ids = np.ascontiguousarray(cand_ids, np.int32)
        cand_dists = np.ascontiguousarray(cand_dists, np.float32)
        ids = np.empty(len(cand_ids), np.int32)
        ds = np.empty(len(cand_ids), np.float32)
        k = self.L.orc_select_neighbors(self.h, _p(cand_ids), _p(cand_dists), len(cand_ids),
                                        max_neighbors, _p(ids), _p(ds))
        return ids[:k].copy(), ds[:k].copy()

    # ---- scan --------------------------------------------------------------------------
    def search_batch(self, queries, ef_search=40, k=10, nthreads=1, with_counters=False):
        q = as_rows(self.elem_type, self.dim, queries)
        nq = len(q)
        ids = np.empty((nq, k), np.int32)
        ds = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.int32)
        c = Counters()
        rc = self.L.orc_search_batch(self.h, _p(q), nq, ef_search, k, _p(ids), _p(ds), _p(cnt),
                                     C.byref(c), nthreads)
        if rc != 0:
            raise ValueError("orc_search_batch rc=%d" % rc)
        return (ids, ds, cnt, c.as_dict()) if with_counters else (ids, ds, cnt)

    def scan_stream(self, query, ef_search=40, iterative=ITER_OFF, max_scan_tuples=20000,
                    max_emit=1 << 20):
        q = as_rows(self.elem_type, self.dim, query)
        max_emit = int(min(max_emit, 10 * max(self.n, 1) + 16))
        tids = np.empty(max_emit, np.int64)
        eids = np.empty(max_emit, np.int32)
        ds = np.empty(max_emit, np.float64)
        c = Counters()
        k = self.L.orc_scan_stream(self.h, _p(q), ef_search, iterative, max_scan_tuples, max_emit,
                                   _p(tids), _p(eids), _p(ds), C.byref(c))
        return tids[:k].copy(), eids[:k].copy(), ds[:k].copy(), c.as_dict()


def flat_topk(elem_type, metric, dim, rows, queries, k, nthreads=1):
    rows = as_rows(elem_type, dim, rows)
    q = as_rows(elem_type, dim, queries)
    ids = np.empty((len(q), k), np.int32)
    ds = np.empty((len(q), k), np.float32)
    lib().orc_flat_topk(elem_type, metric, dim, _p(rows), len(rows), _p(q), len(q), k, _p(ids), _p(ds),
                        nthreads)
    return ids, ds
