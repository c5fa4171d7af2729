"""pgvector_rx_b200 -- B200-native HNSW hot path for pgvector-rx behind a C ABI.

This module is the thin Python binding of libpgvrx_b200.so (include/pgvrx_b200.h) used by the
tests and bench.py.  All computation happens in the CUDA library; there is no CPU fallback and
loading fails loudly when the library has not been built.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PGVRX_LIB selects an alternative build of the same library (e.g. the phase-timing profiling build)
LIB_PATH = os.environ.get("PGVRX_LIB") or os.path.join(_HERE, "libpgvrx_b200.so")

ELEM_F32, ELEM_F16, ELEM_BIT = 0, 1, 2
L2SQ, NEG_IP, L1, HAMMING, JACCARD = 0, 1, 2, 3, 4
ITER_OFF, ITER_RELAXED, ITER_STRICT = 0, 1, 2
OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED = range(6)

# operator classes of the reference (src/types/vector.rs:842-865, halfvec.rs:1050-1073,
# bitvec.rs:210-237): name -> (element type, FUNCTION 1 metric, has FUNCTION 2 norm)
OPCLASSES = {
    "vector_l2_ops": (ELEM_F32, L2SQ, False),
    "vector_ip_ops": (ELEM_F32, NEG_IP, False),
    "vector_cosine_ops": (ELEM_F32, NEG_IP, True),
    "vector_l1_ops": (ELEM_F32, L1, False),
    "halfvec_l2_ops": (ELEM_F16, L2SQ, False),
    "halfvec_ip_ops": (ELEM_F16, NEG_IP, False),
    "halfvec_cosine_ops": (ELEM_F16, NEG_IP, True),
    "halfvec_l1_ops": (ELEM_F16, L1, False),
    "bit_hamming_ops": (ELEM_BIT, HAMMING, False),
    "bit_jaccard_ops": (ELEM_BIT, JACCARD, False),
}

# every symbol include/pgvrx_b200.h declares
EXPORTS = [
    "pgvrx_abi_version", "pgvrx_last_error", "pgvrx_device_count", "pgvrx_normalize",
    "pgvrx_level_from_uniform", "pgvrx_max_level", "pgvrx_row_bytes", "pgvrx_index_create",
    "pgvrx_index_destroy", "pgvrx_index_load_graph", "pgvrx_index_size", "pgvrx_search_batch",
    "pgvrx_search_batch_device", "pgvrx_pages_inspect", "pgvrx_pages_export", "pgvrx_index_build",
    "pgvrx_index_graph_info", "pgvrx_index_export_graph", "pgvrx_index_load_rows", "pgvrx_flat_topk",
    "pgvrx_flat_topk_device", "pgvrx_rerank", "pgvrx_rerank_device", "pgvrx_topk_merge_device", "pgvrx_topk_merge_parts_device",
    "pgvrx_scan_begin", "pgvrx_scan_next", "pgvrx_scan_next_rounds", "pgvrx_scan_end", "pgvrx_flat_tc_scores", "pgvrx_normalize_rows",
    "pgvrx_pages_element_flags", "pgvrx_heap_tid_stream", "pgvrx_index_set_element_flags", "pgvrx_index_insert",
    "pgvrx_index_set_heap_tids", "pgvrx_pages_write", "pgvrx_index_vacuum", "pgvrx_index_element_flags",
    "pgvrx_gather_probe", "pgvrx_group_create", "pgvrx_group_destroy", "pgvrx_group_last_error", "pgvrx_group_size",
    "pgvrx_group_member", "pgvrx_group_peer_access", "pgvrx_group_load_graph", "pgvrx_group_search_batch",
    "pgvrx_group_load_rows", "pgvrx_group_flat_topk", "pgvrx_search_layer_batch", "pgvrx_commit_insert",
]
ELEM_DELETED, ELEM_NO_HEAPTIDS = 1, 2


class PgvrxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pgvrx error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


class IndexDesc(C.Structure):
    _fields_ = [("elem_type", C.c_int32), ("metric", C.c_int32), ("dim", C.c_int32), ("m", C.c_int32),
                ("ef_construction", C.c_int32), ("device", C.c_int32), ("capacity", C.c_int64)]


class PagesInfo(C.Structure):
    _fields_ = [("n_elements", C.c_int64), ("n_upper_rows", C.c_int64), ("n_heap_tids", C.c_int64),
                ("n_deleted", C.c_int64), ("entry", C.c_int64), ("entry_level", C.c_int32), ("dims", C.c_int32),
                ("m", C.c_int32), ("ef_construction", C.c_int32)]


class BuildStats(C.Structure):
    _fields_ = [("n_rounds", C.c_uint64), ("n_searched", C.c_uint64), ("n_conflicts", C.c_uint64),
                ("n_duplicates", C.c_uint64), ("n_big_retries", C.c_uint64), ("n_dist", C.c_uint64),
                ("n_expand0", C.c_uint64), ("n_expand_upper", C.c_uint64), ("n_pair", C.c_uint64),
                ("n_pair_gpu", C.c_uint64), ("n_validation_dist", C.c_uint64), ("n_host_dist", C.c_uint64),
                ("gpu_ms", C.c_double), ("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("host_threads", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class VacuumStats(C.Structure):
    _fields_ = [("n_heap_tids_removed", C.c_uint64), ("n_elements_deleted", C.c_uint64), ("n_elements_repaired", C.c_uint64),
                ("n_rounds", C.c_uint64), ("n_searched", C.c_uint64), ("n_conflicts", C.c_uint64),
                ("kernel_ms", C.c_double), ("total_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SearchStats(C.Structure):
    _fields_ = [("n_dist", C.c_uint64), ("n_expand0", C.c_uint64), ("n_expand_upper", C.c_uint64),
                ("n_visited", C.c_uint64), ("n_retried", C.c_uint64), ("n_launches", C.c_uint32),
                ("kernel_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GatherProbeResult(C.Structure):
    _fields_ = [("bytes_per_launch", C.c_uint64), ("ms_per_launch", C.c_double), ("gb_per_s", C.c_double),
                ("warps_per_sm", C.c_int32), ("grid", C.c_int32), ("row_bytes", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GroupStats(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("local_ms", C.c_double), ("merge_kernel_ms", C.c_float), ("n_devices", C.c_int32),
                ("n_peer_direct", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class FlatStats(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("n_candidates", C.c_uint64), ("n_uncertified", C.c_uint64), ("n_launches", C.c_uint32),
                ("n_splits", C.c_uint32), ("path", C.c_uint32), ("kernel_ms", C.c_float), ("phase_ms", C.c_float * 4)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["phase_ms"] = dict(zip(("operand_split", "candidates", "rerank", "certify"), [float(v) for v in self.phase_ms]))
        return d


_lib = None


def lib():
    """Load the CUDA library.  Raises if it is missing: the product has no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built; run `python -m pgvector_rx_b200.build` (needs nvcc)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    sig = {
        "pgvrx_abi_version": (C.c_int, []),
        "pgvrx_last_error": (C.c_char_p, [vp]),
        "pgvrx_device_count": (C.c_int, []),
        "pgvrx_normalize": (C.c_int, [i32, i32, vp, vp, C.POINTER(C.c_double)]),
        "pgvrx_normalize_rows": (C.c_int, [i32, i32, vp, vp, i64, vp]),
        "pgvrx_level_from_uniform": (i32, [C.c_double, i32]),
        "pgvrx_max_level": (i32, [i32]),
        "pgvrx_row_bytes": (C.c_size_t, [i32, i32]),
        "pgvrx_index_create": (C.c_int, [C.POINTER(IndexDesc), C.POINTER(vp)]),
        "pgvrx_index_destroy": (None, [vp]),
        "pgvrx_index_load_graph": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, i64, i64]),
        "pgvrx_index_size": (i64, [vp]),
        "pgvrx_pages_inspect": (C.c_int, [vp, i64, i32, C.POINTER(PagesInfo)]),
        "pgvrx_pages_export": (C.c_int, [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]),
        "pgvrx_pages_element_flags": (C.c_int, [vp, i64, vp]),
        "pgvrx_heap_tid_stream": (C.c_int, [vp, i64, vp, vp, i64, vp, i64, vp]),
        "pgvrx_pages_write": (C.c_int, [i32, i32, i32, i32, i64, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, i64, C.POINTER(i64)]),
        "pgvrx_index_set_element_flags": (C.c_int, [vp, vp, i64]),
        "pgvrx_index_build": (C.c_int, [vp, vp, i64, vp, vp, i32, C.POINTER(BuildStats)]),
        "pgvrx_index_insert": (C.c_int, [vp, vp, i64, vp, vp, i32, C.POINTER(BuildStats)]),
        "pgvrx_index_set_heap_tids": (C.c_int, [vp, vp, vp, i64]),
        "pgvrx_index_vacuum": (C.c_int, [vp, vp, i64, i32, C.POINTER(VacuumStats)]),
        "pgvrx_index_element_flags": (C.c_int, [vp, vp, i64]),
        "pgvrx_index_graph_info": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]),
        "pgvrx_index_export_graph": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "pgvrx_search_batch": (C.c_int, [vp, vp, i64, i32, i32, vp, vp, vp, C.POINTER(SearchStats)]),
        "pgvrx_search_batch_device": (C.c_int, [vp, vp, i64, i32, i32, vp, vp, vp, C.POINTER(SearchStats), vp]),
        "pgvrx_index_load_rows": (C.c_int, [vp, vp, i64]),
        "pgvrx_flat_topk": (C.c_int, [vp, vp, i64, i32, vp, vp, C.POINTER(FlatStats)]),
        "pgvrx_flat_topk_device": (C.c_int, [vp, vp, i64, i32, i64, i64, i64, vp, vp, C.POINTER(FlatStats), vp]),
        "pgvrx_rerank": (C.c_int, [vp, vp, i64, vp, i32, i32, vp, vp]),
        "pgvrx_rerank_device": (C.c_int, [vp, vp, i64, vp, i32, i32, i64, vp, vp, vp]),
        "pgvrx_topk_merge_device": (C.c_int, [vp, vp, vp, i32, i64, i32, i32, vp, vp, vp]),
        "pgvrx_topk_merge_parts_device": (C.c_int, [vp, vp, vp, i32, i64, i32, i32, vp, vp, vp]),
        "pgvrx_flat_tc_scores": (C.c_int, [vp, vp, i64, vp, vp]),
        "pgvrx_scan_begin": (C.c_int, [vp, vp, i64, i32, i32, i64, C.POINTER(vp)]),
        "pgvrx_scan_next": (C.c_int, [vp, vp, vp, vp, vp, vp, C.POINTER(SearchStats)]),
        "pgvrx_scan_next_rounds": (C.c_int, [vp, vp, i32, vp, vp, vp, vp, C.POINTER(SearchStats)]),
        "pgvrx_scan_end": (None, [vp]),
        "pgvrx_gather_probe": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, C.POINTER(GatherProbeResult)]),
        "pgvrx_group_create": (C.c_int, [C.POINTER(IndexDesc), vp, i32, C.POINTER(vp)]),
        "pgvrx_group_destroy": (None, [vp]),
        "pgvrx_group_last_error": (C.c_char_p, [vp]),
        "pgvrx_group_size": (i32, [vp]),
        "pgvrx_group_member": (vp, [vp, i32]),
        "pgvrx_group_peer_access": (C.c_int, [vp, i32]),
        "pgvrx_group_load_graph": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, i64, i64]),
        "pgvrx_group_search_batch": (C.c_int, [vp, vp, i64, i32, i32, vp, vp, vp, C.POINTER(GroupStats)]),
        "pgvrx_group_load_rows": (C.c_int, [vp, vp, i64]),
        "pgvrx_group_flat_topk": (C.c_int, [vp, vp, i64, i32, vp, vp, C.POINTER(GroupStats)]),
        "pgvrx_search_layer_batch": (C.c_int, [vp, vp, vp, i64, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp]),
        "pgvrx_commit_insert": (C.c_int, [vp, vp, i32, vp, i32, vp, vp, vp, C.POINTER(i64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _elem_dtype(elem):
    return {ELEM_F32: np.float32, ELEM_F16: np.uint16, ELEM_BIT: np.uint8}[elem]


def _rows(elem, dim, a):
    a = np.asarray(a)
    if elem == ELEM_F16 and a.dtype == np.float16:
        a = a.view(np.uint16)
    a = np.ascontiguousarray(a, dtype=_elem_dtype(elem))
    if a.ndim == 1:
        a = a[None, :]
    want = (dim + 7) // 8 if elem == ELEM_BIT else dim
    if a.shape[1] != want:
        raise PgvrxError(ERR_INVALID, "expected %d elements per row, got %d" % (want, a.shape[1]))
    return a


def normalize(elem, dim, row):
    """FUNCTION 2 + normalise, bit-identical to the reference (vector.rs:106-126, halfvec.rs:204-233)."""
    a = _rows(elem, dim, row)
    out = np.zeros_like(a)
    norm = C.c_double()
    rc = lib().pgvrx_normalize(elem, dim, _p(a), _p(out), C.byref(norm))
    if rc != OK:
        raise PgvrxError(rc, lib().pgvrx_last_error(None).decode())
    return norm.value, out[0]


def normalize_rows(elem, dim, rows):
    """FUNCTION 2 + normalise for many rows at once: (norms, unit rows); zero-norm rows come out as zero rows."""
    a = _rows(elem, dim, rows)
    out = np.zeros_like(a)
    norms = np.zeros(len(a), np.float64)
    rc = lib().pgvrx_normalize_rows(elem, dim, _p(a), _p(out), len(a), _p(norms))
    if rc != OK:
        raise PgvrxError(rc, lib().pgvrx_last_error(None).decode())
    return norms, out


def export_pages(pages, elem):
    """Index relation pages (bytes / uint8 array, n * 8192) -> graph dict for HnswIndex.load_graph."""
    buf = np.ascontiguousarray(np.frombuffer(pages, dtype=np.uint8) if not isinstance(pages, np.ndarray) else pages)
    n_pages = buf.size // 8192
    info = PagesInfo()
    L = lib()
    rc = L.pgvrx_pages_inspect(_p(buf), n_pages, elem, C.byref(info))
    if rc != OK:
        raise PgvrxError(rc, L.pgvrx_last_error(None).decode())
    n, m = info.n_elements, info.m
    rb = L.pgvrx_row_bytes(elem, info.dims)
    rows = np.zeros((n, rb), np.uint8)
    level = np.zeros(n, np.uint8)
    adj0 = np.zeros((n, 2 * m), np.int32)
    upper_base = np.zeros(n, np.int32)
    adj_up = np.zeros((max(info.n_upper_rows, 1), m), np.int32)
    tids = np.zeros((n, 10), np.uint64)
    cnt = np.zeros(n, np.uint8)
    rc = L.pgvrx_pages_export(_p(buf), n_pages, elem, _p(rows), _p(level), _p(adj0), _p(upper_base), _p(adj_up),
                              _p(tids), _p(cnt))
    if rc != OK:
        raise PgvrxError(rc, L.pgvrx_last_error(None).decode())
    flags = np.zeros(n, np.uint8)
    rc = L.pgvrx_pages_element_flags(_p(buf), n_pages, _p(flags))
    if rc != OK:
        raise PgvrxError(rc, L.pgvrx_last_error(None).decode())
    return dict(flags=flags, n_deleted=info.n_deleted, rows=rows.view(_elem_dtype(elem)).reshape(n, rb // np.dtype(_elem_dtype(elem)).itemsize), level=level, adj0=adj0, upper_base=upper_base,
                adj_up=adj_up[:info.n_upper_rows], entry=info.entry, n=n, m=m, dims=info.dims,
                ef_construction=info.ef_construction, entry_level=info.entry_level, heap_tids=tids,
                heap_tid_counts=cnt)


def write_pages(elem, dims, m, ef_construction, g, heap_tids=None, heap_tid_counts=None):
    """Graph dict (export_graph() / export_pages() layout) -> the index relation's pages as the reference's flush
    phase writes them (build.rs:545-821).  Returns a uint8 array of n_pages * 8192 bytes."""
    L = lib()
    n = int(g["n"])
    rows = _rows(elem, dims, g["rows"]) if n else np.zeros((0, 1), np.uint8)
    level = np.ascontiguousarray(g["level"], np.uint8)
    adj0 = np.ascontiguousarray(g["adj0"], np.int32)
    upper_base = np.ascontiguousarray(g["upper_base"], np.int32)
    adj_up = np.ascontiguousarray(g["adj_up"], np.int32).reshape(-1, m)
    n_upper = len(adj_up)
    if adj_up.size == 0:
        adj_up = np.zeros((1, m), np.int32)
    t = None if heap_tids is None else np.ascontiguousarray(heap_tids, np.uint64)
    c = None if heap_tid_counts is None else np.ascontiguousarray(heap_tid_counts, np.uint8)
    npg = i64_out = C.c_int64(0)
    args = (elem, dims, m, ef_construction, n, _p(rows), _p(level), _p(adj0), _p(upper_base), _p(adj_up), n_upper,
            int(g["entry"]), _p(t), _p(c))
    rc = L.pgvrx_pages_write(*args, None, 0, C.byref(npg))
    if rc != OK:
        raise PgvrxError(rc, L.pgvrx_last_error(None).decode())
    pages = np.zeros(npg.value * 8192, np.uint8)
    rc = L.pgvrx_pages_write(*args, _p(pages), npg.value, C.byref(i64_out))
    if rc != OK:
        raise PgvrxError(rc, L.pgvrx_last_error(None).decode())
    return pages


def heap_tid_stream(elem_ids, heap_tids, heap_tid_counts):
    """amgettuple's fan-out of scan results to heap TIDs: every element emits its TIDs from the BACK of its list
    (src/index/scan.rs:798 `heaptids.pop()`), elements without a TID are skipped (scan.rs:867).  elem_ids: element ids in
    emission order (negative = padding); heap_tids / heap_tid_counts: the n x 10 / n arrays of export_graph() or
    export_pages().  Returns the TIDs as the executor would receive them."""
    ids = np.ascontiguousarray(np.asarray(elem_ids).ravel(), np.int32)
    tids = np.ascontiguousarray(heap_tids, np.uint64).reshape(-1, 10)
    cnt = np.ascontiguousarray(heap_tid_counts, np.uint8)
    out = np.empty(10 * len(ids), np.uint64)
    k = C.c_int64(0)
    rc = lib().pgvrx_heap_tid_stream(_p(ids), len(ids), _p(tids), _p(cnt), len(cnt), _p(out), len(out), C.byref(k))
    if rc != OK:
        raise PgvrxError(rc, lib().pgvrx_last_error(None).decode())
    return out[:k.value].copy()


class HnswIndex:
    """Host-side handle mirroring `CREATE INDEX ... USING hnsw (col <opclass>) WITH (m, ef_construction)`.

    Session settings mirror the GUCs (src/index/options.rs:81-96): `ef_search` (hnsw.ef_search,
    default 40)."""

    def __init__(self, opclass, dim, m=16, ef_construction=64, capacity=0, device=0):
        if opclass not in OPCLASSES:
            raise PgvrxError(ERR_INVALID, "operator class \"%s\" does not exist for access method \"hnsw\"" % opclass)
        self.opclass = opclass
        self.elem, self.metric, self.normalized = OPCLASSES[opclass]
        self.dim, self.m, self.ef_construction = dim, m, ef_construction
        self.ef_search = 40
        self.L = lib()
        d = IndexDesc(self.elem, self.metric, dim, m, ef_construction, device, capacity)
        h = C.c_void_p()
        rc = self.L.pgvrx_index_create(C.byref(d), C.byref(h))
        if rc != OK:
            raise PgvrxError(rc, self.L.pgvrx_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.pgvrx_index_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc != OK:
            raise PgvrxError(rc, self.L.pgvrx_last_error(self.h).decode())

    def __len__(self):
        return self.L.pgvrx_index_size(self.h)

    def load_graph(self, rows, level, adj0, upper_base, adj_up, entry):
        rows = _rows(self.elem, self.dim, rows)
        level = np.ascontiguousarray(level, np.uint8)
        adj0 = np.ascontiguousarray(adj0, np.int32)
        upper_base = np.ascontiguousarray(upper_base, np.int32)
        adj_up = np.ascontiguousarray(adj_up, np.int32).reshape(-1, self.m)
        self._check(self.L.pgvrx_index_load_graph(self.h, _p(rows), len(rows), _p(level), _p(adj0), _p(upper_base),
                                                  _p(adj_up) if len(adj_up) else None, len(adj_up), int(entry)))

    def vacuum(self, dead_tids, host_threads=0):
        """VACUUM: ambulkdelete (src/index/vacuum.rs:816-842) with the heap's dead TIDs.  Returns statistics."""
        t = np.ascontiguousarray(dead_tids, np.uint64)
        st = VacuumStats()
        self._check(self.L.pgvrx_index_vacuum(self.h, _p(t), len(t), host_threads, C.byref(st)))
        return st.as_dict()

    def element_flags(self):
        f = np.zeros(self.graph_info()["n"], np.uint8)
        self._check(self.L.pgvrx_index_element_flags(self.h, _p(f), len(f)))
        return f

    def set_heap_tids(self, heap_tids, counts):
        """Heap TIDs of a loaded index (the n x 10 / n arrays of export_pages / export_graph); see the header."""
        t = np.ascontiguousarray(heap_tids, np.uint64)
        c = np.ascontiguousarray(counts, np.uint8)
        self._check(self.L.pgvrx_index_set_heap_tids(self.h, _p(t), _p(c), len(c)))

    def set_element_flags(self, flags):
        """Per-element vacuum state (ELEM_DELETED / ELEM_NO_HEAPTIDS, scan.rs:179,866-868); None clears."""
        if flags is None:
            self._check(self.L.pgvrx_index_set_element_flags(self.h, None, 0))
            return
        flags = np.ascontiguousarray(flags, np.uint8)
        self._check(self.L.pgvrx_index_set_element_flags(self.h, _p(flags), len(flags)))

    def build(self, rows, levels, heap_tids=None, host_threads=0):
        """CREATE INDEX: insert rows in order (ambuild, build.rs:400-535).  Rows of cosine opclasses are
        normalised here and zero-norm rows dropped, as build_callback does.  Returns build statistics."""
        rows = _rows(self.elem, self.dim, rows)
        levels = np.ascontiguousarray(levels, np.int32)
        tids = None if heap_tids is None else np.ascontiguousarray(heap_tids, np.uint64)
        if self.normalized:
            norms, unit = normalize_rows(self.elem, self.dim, rows)
            keep = np.nonzero(norms != 0.0)[0]
            rows = np.ascontiguousarray(unit[keep])
            levels = levels[keep]
            tids = keep.astype(np.uint64) if tids is None else tids[keep]
        st = BuildStats()
        self._check(self.L.pgvrx_index_build(self.h, _p(rows), len(rows), _p(levels), _p(tids), host_threads, C.byref(st)))
        return st.as_dict()

    def insert(self, rows, levels, heap_tids=None, host_threads=0):
        """INSERT into a live index: aminsert's rules (src/index/insert.rs:1227-1480) for every row in order.
        Cosine opclasses normalise and drop zero-norm rows like build().  Returns the same statistics as build()."""
        rows = _rows(self.elem, self.dim, rows)
        levels = np.ascontiguousarray(levels, np.int32)
        tids = None if heap_tids is None else np.ascontiguousarray(heap_tids, np.uint64)
        if self.normalized:
            norms, unit = normalize_rows(self.elem, self.dim, rows)
            keep = np.nonzero(norms != 0.0)[0]
            rows = np.ascontiguousarray(unit[keep])
            levels = levels[keep]
            tids = keep.astype(np.uint64) if tids is None else tids[keep]
        st = BuildStats()
        self._check(self.L.pgvrx_index_insert(self.h, _p(rows), len(rows), _p(levels), _p(tids), host_threads, C.byref(st)))
        return st.as_dict()

    def graph_info(self):
        n, nu, e, el = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        self._check(self.L.pgvrx_index_graph_info(self.h, C.byref(n), C.byref(nu), C.byref(e), C.byref(el)))
        return dict(n=n.value, n_upper_rows=nu.value, entry=e.value, entry_level=el.value)

    def export_graph(self):
        gi = self.graph_info()
        n, m = gi["n"], self.m
        rb = self.L.pgvrx_row_bytes(self.elem, self.dim)
        rows = np.zeros((n, rb), np.uint8)
        level = np.zeros(n, np.uint8)
        adj0 = np.zeros((n, 2 * m), np.int32)
        upper_base = np.zeros(n, np.int32)
        adj_up = np.zeros((max(gi["n_upper_rows"], 1), m), np.int32)
        tids = np.zeros((n, 10), np.uint64)
        cnt = np.zeros(n, np.uint8)
        self._check(self.L.pgvrx_index_export_graph(self.h, _p(rows), _p(level), _p(adj0), _p(upper_base), _p(adj_up),
                                                    _p(tids), _p(cnt)))
        return dict(rows=rows.view(_elem_dtype(self.elem)).reshape(n, rb // np.dtype(_elem_dtype(self.elem)).itemsize),
                    level=level, adj0=adj0, upper_base=upper_base, adj_up=adj_up[:gi["n_upper_rows"]],
                    entry=gi["entry"], n=n, heap_tids=tids, heap_tid_counts=cnt)

    def prepare_queries(self, queries, prepared=False):
        """Query-side half of the opclass: normalise when FUNCTION 2 exists (scan.rs:749-754).  prepared=True: the
        caller already did (normalising a unit vector again would change its last bits)."""
        q = _rows(self.elem, self.dim, queries)
        if self.normalized and not prepared:
            q = normalize_rows(self.elem, self.dim, q)[1]
        return q

    def load_graph_stored(self, rows, level, adj0, upper_base, adj_up, entry):
        """load_graph under the name bench.py uses to stress that `rows` are in stored form (already normalised)."""
        self.load_graph(rows, level, adj0, upper_base, adj_up, entry)

    def search(self, queries, k=10, ef_search=None, with_stats=False, prepared=False):
        """Batched index scan (hnsw.iterative_scan = off): k nearest element ids per query."""
        q = self.prepare_queries(queries, prepared)
        ef = self.ef_search if ef_search is None else ef_search
        nq = len(q)
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        counts = np.empty(nq, np.int32)
        st = SearchStats()
        self._check(self.L.pgvrx_search_batch(self.h, _p(q), nq, ef, k, _p(ids), _p(dists), _p(counts),
                                              C.byref(st) if with_stats else None))
        return (ids, dists, counts, st.as_dict()) if with_stats else (ids, dists, counts)

    def search_device(self, d_queries_ptr, nq, k, ef_search, d_ids_ptr, d_dists_ptr, d_counts_ptr, stream=0, stats=None):
        """Device-pointer variant (queries already in HBM); asynchronous on `stream` unless stats is given."""
        self._check(self.L.pgvrx_search_batch_device(self.h, d_queries_ptr, nq, ef_search, k, d_ids_ptr, d_dists_ptr,
                                                     d_counts_ptr, C.byref(stats) if stats is not None else None,
                                                     C.c_void_p(stream) if stream else None))

    # ---- exact paths (sequential scan ordered by the distance operator: the reference's ground truth) ----
    def load_rows(self, rows):
        """Vector table only (no graph): enough for flat_topk / rerank.  Cosine opclasses normalise the
        rows like the build does (zero-norm rows are kept as zero rows here so that ids stay row numbers)."""
        rows = _rows(self.elem, self.dim, rows)
        if self.normalized:
            rows = normalize_rows(self.elem, self.dim, rows)[1]
        self._check(self.L.pgvrx_index_load_rows(self.h, _p(rows), len(rows)))

    def load_rows_stored(self, rows):
        """load_rows for rows that are already in stored form (normalised where the opclass normalises)."""
        rows = _rows(self.elem, self.dim, rows)
        self._check(self.L.pgvrx_index_load_rows(self.h, _p(rows), len(rows)))

    def flat_topk(self, queries, k=10, with_stats=False, prepared=False):
        """Exact k nearest rows per query, ordered by (distance, id)."""
        q = self.prepare_queries(queries, prepared)
        nq = len(q)
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        st = FlatStats()
        self._check(self.L.pgvrx_flat_topk(self.h, _p(q), nq, k, _p(ids), _p(dists), C.byref(st) if with_stats else None))
        return (ids, dists, st.as_dict()) if with_stats else (ids, dists)

    def flat_topk_device(self, d_queries_ptr, nq, k, d_ids_ptr, d_dists_ptr, row_begin=0, row_end=-1, id_base=0, stream=0,
                         stats=None):
        self._check(self.L.pgvrx_flat_topk_device(self.h, d_queries_ptr, nq, k, row_begin, row_end, id_base, d_ids_ptr,
                                                  d_dists_ptr, C.byref(stats) if stats is not None else None,
                                                  C.c_void_p(stream) if stream else None))

    def flat_tc_scores(self, queries):
        """Approximate distances of the tensor-core path and its per-query certification bound (diagnostic)."""
        q = self.prepare_queries(queries)
        scores = np.empty((len(q), len(self)), np.float32)
        eps = np.empty(len(q), np.float32)
        self._check(self.L.pgvrx_flat_tc_scores(self.h, _p(q), len(q), _p(scores), _p(eps)))
        return scores, eps

    def rerank(self, queries, cand_ids, k=10, prepared=False):
        """Exact distances of each query's candidate ids (-1 = none); the k nearest by (distance, id)."""
        q = self.prepare_queries(queries, prepared)
        cand = np.ascontiguousarray(cand_ids, np.int32)
        nq, nc = cand.shape
        if nq != len(q):
            raise PgvrxError(ERR_INVALID, "one candidate list per query expected")
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        self._check(self.L.pgvrx_rerank(self.h, _p(q), nq, _p(cand), nc, k, _p(ids), _p(dists)))
        return ids, dists

    def topk_merge_device(self, d_part_ids_ptr, d_part_dists_ptr, n_parts, nq, k_in, k_out, d_out_ids_ptr, d_out_dists_ptr,
                          stream=0):
        self._check(self.L.pgvrx_topk_merge_device(self.h, d_part_ids_ptr, d_part_dists_ptr, n_parts, nq, k_in, k_out,
                                                   d_out_ids_ptr, d_out_dists_ptr, C.c_void_p(stream) if stream else None))

    def topk_merge_parts_device(self, part_ids_ptrs, part_dists_ptrs, nq, k_in, k_out, d_out_ids_ptr, d_out_dists_ptr, stream=0):
        """K6 over per-part device pointers (lists that are not contiguous with each other)."""
        n = len(part_ids_ptrs)
        pi = (C.c_void_p * n)(*part_ids_ptrs)
        pd = (C.c_void_p * n)(*part_dists_ptrs)
        self._check(self.L.pgvrx_topk_merge_parts_device(self.h, pi, pd, n, nq, k_in, k_out, d_out_ids_ptr, d_out_dists_ptr,
                                                         C.c_void_p(stream) if stream else None))

    # ---- the narrow build seam: searches here, heuristic + commit on the caller's side ----
    def search_layer_batch(self, ep_ids, ep_dists, ep_counts, ef, layer, queries=None, query_ids=None, exp_cap=0):
        """search_layer(query, entry_points, ef, layer) for a batch (graph/mod.rs:161-255).  Returns (ids, dists, counts)
        and, with exp_cap > 0, also (expanded ids, expanded counts)."""
        ep_ids = np.ascontiguousarray(ep_ids, np.int32)
        ep_dists = np.ascontiguousarray(ep_dists, np.float32)
        ep_counts = np.ascontiguousarray(ep_counts, np.int32)
        nq, pitch = ep_ids.shape
        q = None if queries is None else _rows(self.elem, self.dim, queries)
        qi = None if query_ids is None else np.ascontiguousarray(query_ids, np.int32)
        ids = np.empty((nq, ef), np.int32)
        dists = np.empty((nq, ef), np.float32)
        counts = np.empty(nq, np.int32)
        ex = np.empty((nq, max(exp_cap, 1)), np.int32)
        nex = np.empty(nq, np.int32)
        self._check(self.L.pgvrx_search_layer_batch(self.h, _p(q), _p(qi), nq, _p(ep_ids), _p(ep_dists), _p(ep_counts), pitch, ef, layer,
                                                    _p(ids), _p(dists), _p(counts), _p(ex) if exp_cap else None, exp_cap,
                                                    _p(nex) if exp_cap else None))
        return (ids, dists, counts, ex, nex) if exp_cap else (ids, dists, counts)

    def commit_insert(self, row, level, lists, mod_nodes=(), mod_layers=(), mod_lists=()):
        """One new element as the caller decided it (see the header).  Returns its element id."""
        r = _rows(self.elem, self.dim, row)
        lists = np.ascontiguousarray(lists, np.int32)
        mn = np.ascontiguousarray(mod_nodes, np.int32)
        ml = np.ascontiguousarray(mod_layers, np.int32)
        mls = np.ascontiguousarray(mod_lists, np.int32).ravel()
        new_id = C.c_int64(-1)
        self._check(self.L.pgvrx_commit_insert(self.h, _p(r), int(level), _p(lists), len(mn), _p(mn) if len(mn) else None,
                                               _p(ml) if len(mn) else None, _p(mls) if len(mn) else None, C.byref(new_id)))
        return new_id.value

    def gather_probe(self, mode=1, warps_per_sm=0, rows_per_step=32, depth=1, steps=256, consume=1):
        """Diagnostic: bytes/s of random row gathers with the scan kernel's TMA mechanism (see the header)."""
        r = GatherProbeResult()
        self._check(self.L.pgvrx_gather_probe(self.h, mode, warps_per_sm, rows_per_step, depth, steps, consume, C.byref(r)))
        return r.as_dict()

    # ---- iterative scan (hnsw.iterative_scan, hnsw.max_scan_tuples: options.rs:85-91) ----
    def scan_begin(self, queries, ef_search=None, iterative_scan=ITER_RELAXED, max_scan_tuples=20000, prepared=False):
        return IterativeScan(self, queries, self.ef_search if ef_search is None else ef_search, iterative_scan,
                             max_scan_tuples, prepared)


class HnswGroup:
    """Several GPUs of one box behind one handle (pgvrx_group_*): query-sharded scans over a replicated graph, exact flat
    kNN over row shards with the merge reading the peers' lists over NVLink.  Rows / queries are taken in STORED form
    (already normalised where the opclass normalises)."""

    def __init__(self, opclass, dim, devices, m=16, ef_construction=64, capacity=0):
        if opclass not in OPCLASSES:
            raise PgvrxError(ERR_INVALID, "operator class \"%s\" does not exist for access method \"hnsw\"" % opclass)
        self.elem, self.metric, self.normalized = OPCLASSES[opclass]
        self.dim, self.m = dim, m
        self.L = lib()
        d = IndexDesc(self.elem, self.metric, dim, m, ef_construction, 0, capacity)
        devs = np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        rc = self.L.pgvrx_group_create(C.byref(d), _p(devs), len(devs), C.byref(h))
        if rc != OK:
            raise PgvrxError(rc, self.L.pgvrx_group_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.pgvrx_group_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc != OK:
            raise PgvrxError(rc, self.L.pgvrx_group_last_error(self.h).decode())

    def __len__(self):
        return self.L.pgvrx_group_size(self.h)

    def peer_access(self):
        return [bool(self.L.pgvrx_group_peer_access(self.h, i)) for i in range(len(self))]

    def load_graph(self, rows, level, adj0, upper_base, adj_up, entry):
        rows = _rows(self.elem, self.dim, rows)
        level = np.ascontiguousarray(level, np.uint8)
        adj0 = np.ascontiguousarray(adj0, np.int32)
        upper_base = np.ascontiguousarray(upper_base, np.int32)
        adj_up = np.ascontiguousarray(adj_up, np.int32).reshape(-1, self.m)
        self._check(self.L.pgvrx_group_load_graph(self.h, _p(rows), len(rows), _p(level), _p(adj0), _p(upper_base),
                                                  _p(adj_up) if len(adj_up) else None, len(adj_up), int(entry)))

    def search(self, queries, k=10, ef_search=40, with_stats=False, out=None):
        q = _rows(self.elem, self.dim, queries)
        nq = len(q)
        ids, dists, counts = out if out is not None else (np.empty((nq, k), np.int32), np.empty((nq, k), np.float32), np.empty(nq, np.int32))
        st = GroupStats()
        self._check(self.L.pgvrx_group_search_batch(self.h, _p(q), nq, ef_search, k, _p(ids), _p(dists), _p(counts), C.byref(st)))
        return (ids, dists, counts, st.as_dict()) if with_stats else (ids, dists, counts)

    def load_rows(self, rows):
        rows = _rows(self.elem, self.dim, rows)
        self._check(self.L.pgvrx_group_load_rows(self.h, _p(rows), len(rows)))

    def flat_topk(self, queries, k=10, with_stats=False):
        q = _rows(self.elem, self.dim, queries)
        nq = len(q)
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        st = GroupStats()
        self._check(self.L.pgvrx_group_flat_topk(self.h, _p(q), nq, k, _p(ids), _p(dists), C.byref(st)))
        return (ids, dists, st.as_dict()) if with_stats else (ids, dists)


class IterativeScan:
    """amrescan .. amendscan for a batch of queries.  `next()` returns, per query, the next block of
    elements in amgettuple's emission order; `stream(limit)` drives rounds until every query has
    `limit` elements or is exhausted."""

    def __init__(self, index, queries, ef_search, mode, max_scan_tuples, prepared=False):
        self.ix, self.ef = index, ef_search
        q = index.prepare_queries(queries, prepared)
        self.nq = len(q)
        h = C.c_void_p()
        index._check(index.L.pgvrx_scan_begin(index.h, _p(q), self.nq, ef_search, mode, max_scan_tuples, C.byref(h)))
        self.h = h
        self.totals = dict(n_dist=0, n_expand0=0, n_expand_upper=0, n_visited=0, kernel_ms=0.0, rounds=0)

    def next(self, want_more=None, rounds=None):
        """One refill per query (rounds=None: arrays [nq, ef] / [nq]) or `rounds` refills in one launch
        (arrays [nq, rounds, ef] / [nq, rounds]: batch r of a query follows batch r - 1 in its stream)."""
        r = 1 if rounds is None else int(rounds)
        ids = np.empty((self.nq, r, self.ef), np.int32)
        dists = np.empty((self.nq, r, self.ef), np.float32)
        counts = np.empty((self.nq, r), np.int32)
        done = np.empty(self.nq, np.uint8)
        wm = None if want_more is None else np.ascontiguousarray(want_more, np.uint8)
        st = SearchStats()
        self.ix._check(self.ix.L.pgvrx_scan_next_rounds(self.h, _p(wm), r, _p(ids), _p(dists), _p(counts), _p(done), C.byref(st)))
        for k in ("n_dist", "n_expand0", "n_expand_upper", "n_visited", "kernel_ms"):
            self.totals[k] += getattr(st, k)
        self.totals["rounds"] += r
        if rounds is None:
            return ids[:, 0], dists[:, 0], counts[:, 0], done.astype(bool)
        return ids, dists, counts, done.astype(bool)

    def stream(self, limit):
        """Per query: (element ids, distances) in emission order, at most `limit` (None = until exhausted)."""
        out_i = [[] for _ in range(self.nq)]
        out_d = [[] for _ in range(self.nq)]
        have = np.zeros(self.nq, np.int64)
        done = np.zeros(self.nq, bool)
        while True:
            want = ~done if limit is None else (~done & (have < limit))
            if not want.any():
                break
            ids, dists, counts, d = self.next(want.astype(np.uint8))
            for q in np.nonzero(want)[0]:
                c = counts[q]
                out_i[q].append(ids[q, :c])
                out_d[q].append(dists[q, :c])
                have[q] += c
            done |= d
        cat = lambda parts, dt: np.concatenate(parts) if parts else np.zeros(0, dt)
        res_i = [cat(p, np.int32)[:limit] for p in out_i]
        res_d = [cat(p, np.float32)[:limit] for p in out_d]
        return res_i, res_d

    def close(self):
        if getattr(self, "h", None):
            self.ix.L.pgvrx_scan_end(self.h)
            self.h = None

    __del__ = close
