"""a19: the product's page exporter against pages written by the oracle exactly as the reference's
flush phase does (build.rs:545-821).  No PostgreSQL exists in this image, so the fixture is
synthesised; the exporter and the writer are independent implementations of the same layout."""
import os
import struct

import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import make_rows, oracle_index, vacuumed_oracle
from pgvector_rx_b200 import datasets as D


def roundtrip(opclass, kind, n, dim, m, efc, levels=None, seed=5):
    elem = P.OPCLASSES[opclass][0]
    rows = make_rows(kind, n, dim, seed)
    o = oracle_index(opclass, dim, rows, m=m, efc=efc, levels=levels)
    pages = o.write_pages()
    g = P.export_pages(pages, elem)
    want = o.export_graph()
    assert g["n"] == want["n"] and g["m"] == m and g["dims"] == dim and g["ef_construction"] == efc
    assert g["entry"] == want["entry"]
    if want["n"]:
        assert g["entry_level"] == want["level"][want["entry"]]
    for k in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(g[k], want[k]), k
    assert np.array_equal(g["rows"].view(np.uint8), want["rows"].view(np.uint8))
    for i in range(want["n"]):
        t = o.heap_tids(i)
        assert g["heap_tid_counts"][i] == len(t)
        got = [((int(v) >> 16), (int(v) & 0xffff) - 1) for v in g["heap_tids"][i, :len(t)]]
        assert got == [(int(x) >> 16, int(x) & 0xffff) for x in t]
    return pages, g


@pytest.mark.parametrize("opclass,kind,dim", [("vector_l2_ops", "f32", 128), ("vector_cosine_ops", "f32", 7),
                                              ("halfvec_ip_ops", "f16", 33), ("bit_hamming_ops", "bit", 52),
                                              ("bit_jaccard_ops", "bit", 1024)])
def test_export_matches_builder(opclass, kind, dim):
    roundtrip(opclass, kind, 700, dim, 8, 32)


def test_export_page_layout_edges():
    # many elements per page / neighbour tuple forced onto the next page (combined > HNSW_MAX_SIZE,
    # build.rs:653-657) / very high levels
    roundtrip("vector_l2_ops", "f32", 300, 3, 4, 16)
    pages, g = roundtrip("vector_l2_ops", "f32", 40, 2000, 16, 64)
    assert len(pages) // 8192 >= 2 * 40
    lv = np.zeros(60, np.int32)
    lv[[3, 17, 40]] = [9, 40, 2]
    roundtrip("vector_l2_ops", "f32", 60, 16, 32, 64, levels=lv)


def test_export_duplicates_and_empty():
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    o.build(np.tile(np.array([1, 2, 3], np.float32), (25, 1)), np.zeros(25, np.int32))
    g = P.export_pages(o.write_pages(), P.ELEM_F32)
    assert g["n"] == 3 and sorted(g["heap_tid_counts"].tolist()) == [5, 10, 10]
    e = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    g = P.export_pages(e.write_pages(), P.ELEM_F32)
    assert g["n"] == 0 and g["entry"] == -1                     # meta + head page only (build.rs:580-589)


def test_export_rejects_corrupt_relations():
    o = oracle_index("vector_l2_ops", 8, make_rows("f32", 50, 8, 1), m=4, efc=16)
    pages = o.write_pages()
    bad = pages.copy()
    bad[24:28] = 0                                               # magic number
    with pytest.raises(P.PgvrxError) as e:
        P.export_pages(bad, P.ELEM_F32)
    assert "hnsw index is not valid" in str(e.value)            # scan.rs:134-137
    with pytest.raises(P.PgvrxError):
        P.export_pages(pages, P.ELEM_F16)                       # value size does not match the type
    with pytest.raises(P.PgvrxError):
        P.export_pages(pages[:8192], P.ELEM_F32)
    # stale neighbour tuple (version mismatch) => that element simply has no lists (scan.rs:262)
    g0 = P.export_pages(pages, P.ELEM_F32)
    stale = pages.copy()
    page1 = stale[8192:16384]
    lp = struct.unpack_from("<I", page1, 24 + 4)[0]              # item 2 of the head page = first neighbour tuple
    off = lp & 0x7fff
    assert page1[off] == 2
    page1[off + 1] = 7
    g1 = P.export_pages(stale, P.ELEM_F32)
    assert (g1["adj0"][0] == -1).all() and np.array_equal(g1["adj0"][1:], g0["adj0"][1:])


def test_export_vacuumed_index():
    o, deleted, no_tid, stale = vacuumed_oracle("vector_l2_ops", "f32", 600, 12, 6, 24)
    want = o.export_graph()
    g = P.export_pages(o.write_pages(), P.ELEM_F32)
    assert g["n"] == want["n"] and g["n_deleted"] == len(deleted) and g["entry"] == want["entry"]
    flags = o.element_flags()
    assert np.array_equal(g["flags"], flags)
    assert (flags[deleted] == (P.ELEM_DELETED | P.ELEM_NO_HEAPTIDS)).all() and (flags[no_tid] == P.ELEM_NO_HEAPTIDS).all()
    assert (g["rows"][deleted] == 0).all() and (g["adj0"][deleted] == -1).all() and (g["heap_tid_counts"][deleted] == 0).all()
    # stale neighbour tuples read as empty lists; every other element is exported unchanged
    assert (g["adj0"][stale] == -1).all()
    live = np.setdiff1d(np.arange(o.n), np.concatenate([deleted, stale]))
    assert np.array_equal(g["adj0"][live], want["adj0"][live]) and np.array_equal(g["rows"][live], want["rows"][live])
    assert np.array_equal(g["level"], want["level"])


def test_oracle_scan_semantics_on_vacuumed_index():
    # scan.rs:179: deleted elements are visited but never returned; scan.rs:866-868: elements without heap TIDs
    # are searched through but emit nothing; a deleted entry point returns no rows (scan.rs:475-486)
    o, deleted, no_tid, stale = vacuumed_oracle("vector_l2_ops", "f32", 600, 12, 6, 24)
    q = make_rows("f32", 40, 12, 3)
    ids, dists, counts = o.search_batch(q, ef_search=40, k=10)
    got = set(ids[ids >= 0].tolist())
    assert not (got & set(deleted.tolist())) and not (got & set(no_tid.tolist()))
    assert (counts > 0).all()
    for mode in (O.ITER_RELAXED, O.ITER_STRICT):
        tids, eids, ds, _ = o.scan_stream(q[0], ef_search=10, iterative=mode)
        assert not (set(eids.tolist()) & (set(deleted.tolist()) | set(no_tid.tolist())))
    o2, *_ = vacuumed_oracle("vector_l2_ops", "f32", 600, 12, 6, 24, delete_entry=True)
    ids, dists, counts = o2.search_batch(q, ef_search=40, k=10)
    assert (counts == 0).all() and (ids == -1).all()


# ---- the product's page WRITER (pgvrx_pages_write): byte-identical to the oracle's independent writer -----------------
def fixture_tids(o):
    """Heap TIDs in the product's (block << 16 | offset) form, under the oracle fixture's convention
    (heap TID t -> block t >> 16, offset (t & 0xffff) + 1)."""
    n = o.n
    tids = np.full((n, 10), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    cnt = np.zeros(n, np.uint8)
    for i in range(n):
        t = o.heap_tids(i)
        cnt[i] = len(t)
        for j, v in enumerate(t):
            tids[i, j] = ((int(v) >> 16) << 16) | ((int(v) & 0xffff) + 1)
    return tids, cnt


def write_matches(opclass, kind, n, dim, m, efc, levels=None, seed=5):
    elem = P.OPCLASSES[opclass][0]
    o = oracle_index(opclass, dim, make_rows(kind, n, dim, seed), m=m, efc=efc, levels=levels)
    want = o.write_pages()
    tids, cnt = fixture_tids(o)
    got = P.write_pages(elem, dim, m, efc, o.export_graph(), tids, cnt)
    assert len(got) == len(want), (len(got) // 8192, len(want) // 8192)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, "first difference in page %d at byte %d" % (bad[0] // 8192, bad[0] % 8192)
    # and back through the exporter
    g = P.export_pages(got, elem)
    w = o.export_graph()
    for k in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(g[k], w[k]), k
    assert g["entry"] == w["entry"] and np.array_equal(g["heap_tid_counts"], cnt) and np.array_equal(g["heap_tids"], tids)
    return got


@pytest.mark.parametrize("opclass,kind,dim", [("vector_l2_ops", "f32", 128), ("vector_cosine_ops", "f32", 7),
                                              ("halfvec_ip_ops", "f16", 33), ("bit_hamming_ops", "bit", 52),
                                              ("bit_jaccard_ops", "bit", 1024)])
def test_writer_matches_reference_flush(opclass, kind, dim):
    write_matches(opclass, kind, 700, dim, 8, 32)


def test_writer_page_layout_edges():
    write_matches("vector_l2_ops", "f32", 300, 3, 4, 16)                 # many elements per page
    pages = write_matches("vector_l2_ops", "f32", 40, 2000, 16, 64)      # neighbour tuple forced onto the next page
    assert len(pages) // 8192 >= 2 * 40
    lv = np.zeros(60, np.int32)
    lv[[3, 17, 40]] = [9, 40, 2]
    write_matches("vector_l2_ops", "f32", 60, 16, 32, 64, levels=lv)     # very high levels
    # duplicates (several heap TIDs per element) and the empty index
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    o.build(np.tile(np.array([1, 2, 3], np.float32), (25, 1)), np.zeros(25, np.int32))
    tids, cnt = fixture_tids(o)
    assert np.array_equal(P.write_pages(P.ELEM_F32, 3, 16, 64, o.export_graph(), tids, cnt), o.write_pages())
    e = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    empty = dict(n=0, rows=np.zeros((0, 3), np.float32), level=np.zeros(0, np.uint8), adj0=np.zeros((0, 32), np.int32),
                 upper_base=np.zeros(0, np.int32), adj_up=np.zeros((0, 16), np.int32), entry=-1)
    assert np.array_equal(P.write_pages(P.ELEM_F32, 3, 16, 64, empty), e.write_pages())


def synthetic_graph(n, dim, m, seed):
    """Any well-formed graph will do for the page format: random levels, random dense-prefix lists."""
    rng = np.random.default_rng(seed)
    level = np.minimum(rng.geometric(0.7, n) - 1, 6).astype(np.uint8)
    upper_base = np.full(n, -1, np.int32)
    upper_base[level > 0] = (np.cumsum(level) - level)[level > 0]
    n_up = int(level.sum())
    adj0 = rng.integers(0, n, (n, 2 * m)).astype(np.int32)
    adj0[np.arange(2 * m)[None, :] >= rng.integers(0, 2 * m + 1, n)[:, None]] = -1
    adj_up = rng.integers(0, n, (n_up, m)).astype(np.int32)
    adj_up[np.arange(m)[None, :] >= rng.integers(0, m + 1, n_up)[:, None]] = -1
    return dict(n=n, rows=rng.standard_normal((n, dim)).astype(np.float32), level=level, adj0=adj0, upper_base=upper_base,
                adj_up=adj_up, entry=int(np.argmax(level)))


def test_writer_and_exporter_on_their_threaded_paths():
    # above 32k elements / 1k pages both directions decode on several host threads: the bytes must not depend on it
    n, dim, m = 60_000, 24, 4                            # below 65535: the oracle fixture's TID convention wraps there
    g = synthetic_graph(n, dim, m, 11)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, dim, m=m, ef_construction=16)
    o.import_graph(g)
    want = o.write_pages()
    got = P.write_pages(P.ELEM_F32, dim, m, 16, g)
    assert len(want) // 8192 > 1024 and np.array_equal(got, want)
    e = P.export_pages(got, P.ELEM_F32)
    for k in ("rows", "level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(e[k], g[k]), k
    assert e["entry"] == g["entry"] and e["n"] == n and not e["flags"].any()
    # the writer's default heap TIDs (none given) stay valid and distinct past 65535 rows
    big = synthetic_graph(70_000, 2, 2, 12)
    eb = P.export_pages(P.write_pages(P.ELEM_F32, 2, 2, 8, big), P.ELEM_F32)
    assert (eb["heap_tid_counts"] == 1).all() and not eb["flags"].any()
    assert len(np.unique(eb["heap_tids"][:, 0])) == 70_000 and ((eb["heap_tids"][:, 0] & 0xffff) >= 1).all()
    # a neighbour id out of range is refused from whichever thread meets it
    bad = dict(g, adj0=g["adj0"].copy())
    bad["adj0"][n - 5, 0] = n
    with pytest.raises(P.PgvrxError):
        P.write_pages(P.ELEM_F32, dim, m, 16, bad)
    # and so is a neighbour TID that points at no element tuple, on the way back
    torn = got.copy()
    info_off = 8192 * (len(torn) // 8192 - 1)            # last page: zero its first line pointer's tuple type
    lp = int(np.frombuffer(torn[info_off + 24: info_off + 28].tobytes(), np.uint32)[0])
    torn[info_off + (lp & 0x7fff)] = 7                   # neither an element nor a neighbour tuple
    try:
        P.export_pages(torn, P.ELEM_F32)                 # accepted only if nothing pointed at that tuple
    except P.PgvrxError as ex:
        assert "does not reference an element tuple" in str(ex) or "entry point" in str(ex)


def test_writer_rejects_heap_tids_a_reader_would_drop():
    # a TID with offset 0 (e.g. a bare row number used as a TID) reads back as "no heap TID" (scan.rs:203-209)
    o = oracle_index("vector_l2_ops", 4, make_rows("f32", 50, 4, 2), m=4, efc=16)
    tids, cnt = fixture_tids(o)
    tids[7, 0] = np.uint64(3 << 16)
    with pytest.raises(P.PgvrxError) as e:
        P.write_pages(P.ELEM_F32, 4, 4, 16, o.export_graph(), tids, cnt)
    assert "offset >= 1" in str(e.value)


def test_writer_rejects_elements_a_vacuum_emptied():
    # an element without heap TIDs is a vacuum's product (deleted = 1, lists cleared, versions bumped, vacuum.rs:655-790);
    # the writer lays out the CREATE INDEX flush only and must not turn it into a live tuple without TIDs
    o = oracle_index("vector_l2_ops", 4, make_rows("f32", 50, 4, 2), m=4, efc=16)
    tids, cnt = fixture_tids(o)
    cnt[9] = 0
    tids[9] = np.uint64(0xFFFFFFFFFFFFFFFF)
    with pytest.raises(P.PgvrxError) as e:
        P.write_pages(P.ELEM_F32, 4, 4, 16, o.export_graph(), tids, cnt)
    assert "CREATE INDEX flush only" in str(e.value)


def test_writer_rejects_bad_input():
    o = oracle_index("vector_l2_ops", 8, make_rows("f32", 50, 8, 1), m=4, efc=16)
    g = o.export_graph()
    bad = dict(g)
    bad["adj0"] = g["adj0"].copy()
    bad["adj0"][0, 0] = 50
    with pytest.raises(P.PgvrxError):
        P.write_pages(P.ELEM_F32, 8, 4, 16, bad)
    with pytest.raises(P.PgvrxError):
        P.write_pages(P.ELEM_F32, 8, 1, 16, g)                          # m below the reloption minimum
    with pytest.raises(P.PgvrxError) as e:
        P.write_pages(P.ELEM_F32, 2100, 4, 16, dict(g, rows=np.zeros((50, 2100), np.float32)))
    assert "index tuple too large" in str(e.value)                       # build.rs:612-614


def test_export_after_a_real_vacuum():
    # the relation the oracle's ambulkdelete restatement leaves behind (repaired lists, deleted elements with zeroed values
    # and bumped versions) goes through the exporter unchanged
    o = oracle_index("vector_l2_ops", 10, make_rows("f32", 900, 10, 3), m=5, efc=20)
    dead = np.random.default_rng(4).permutation(900)[:300]
    assert o.bulk_delete(dead) == 300
    want = o.export_graph()
    g = P.export_pages(o.write_pages(), P.ELEM_F32)
    assert g["n_deleted"] == 300 and np.array_equal(g["flags"], o.element_flags())
    for k in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(g[k], want[k]), k
    assert g["entry"] == want["entry"] and np.array_equal(g["rows"], want["rows"])
    gone = np.nonzero(g["flags"] & P.ELEM_DELETED)[0]
    assert sorted(gone.tolist()) == sorted(dead.tolist())
    live_lists = g["adj0"][g["flags"] == 0]
    assert not np.isin(live_lists[live_lists >= 0], gone).any()         # repair removed every pointer to a deleted element


# ---- amgettuple's heap TID fan-out (pgvrx_heap_tid_stream) -----------------------------------------------------------
def test_heap_tid_stream_matches_amgettuple_order():
    # duplicates: 25 copies of one row and 3 of another -> elements holding 10 / 10 / 5 / 3 heap TIDs; the oracle's
    # amgettuple restatement emits each element's TIDs from the back of its list (scan.rs:798)
    rows = np.concatenate([np.tile(np.array([1, 2, 3], np.float32), (25, 1)), np.tile(np.array([4, 4, 4], np.float32), (3, 1)),
                           make_rows("f32", 40, 3, 3)])
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    o.build(rows, np.zeros(len(rows), np.int32))
    g = P.export_pages(o.write_pages(), P.ELEM_F32)
    assert sorted(g["heap_tid_counts"].tolist(), reverse=True)[:4] == [10, 10, 5, 3]
    for q in (rows[0], rows[26], rows[40]):
        tids, eids, _, _ = o.scan_stream(q, ef_search=40)
        elems = eids[np.concatenate([[True], eids[1:] != eids[:-1]])]          # one entry per emitted element
        got = P.heap_tid_stream(np.concatenate([elems, [-1, -1]]), g["heap_tids"], g["heap_tid_counts"])
        want = [((int(t) >> 16) << 16) | ((int(t) & 0xffff) + 1) for t in tids]  # the fixture's TID form (fixture_tids)
        assert got.tolist() == want
    # an element without heap TIDs contributes nothing; bad ids and short buffers are reported, not written past
    cnt = g["heap_tid_counts"].copy()
    cnt[0] = 0
    assert len(P.heap_tid_stream([0, 1], g["heap_tids"], cnt)) == cnt[1]
    with pytest.raises(P.PgvrxError):
        P.heap_tid_stream([g["n"]], g["heap_tids"], g["heap_tid_counts"])
    import ctypes as C
    ids = np.array([0, 1], np.int32)
    out = np.full(4, 7, np.uint64)
    k = C.c_int64(0)
    assert P.lib().pgvrx_heap_tid_stream(P._p(ids), 2, P._p(g["heap_tids"]), P._p(g["heap_tid_counts"]), g["n"], P._p(out), 3, C.byref(k)) == 0
    assert k.value == int(g["heap_tid_counts"][0]) + int(g["heap_tid_counts"][1]) > 3 and out[3] == 7


# ---- the on-disk format, frozen ---------------------------------------------------------------------------------------
def test_relation_bytes_match_the_frozen_digests():
    # tests/golden/pages_sha256.json (made by tests/golden/make_pages_golden.py): both page writers -- the oracle's and
    # the product's pgvrx_pages_write -- must still produce those bytes, so the two cannot drift together
    import hashlib
    import importlib.util
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_pages_golden", os.path.join(here, "make_pages_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    want = json.load(open(os.path.join(here, "pages_sha256.json")))
    assert set(want) == set(G.CASES) | {"vector_l2_vacuumed"}
    for name, (opclass, kind, n, dim, m, efc, seed) in G.CASES.items():
        o = G.relation(name)
        pages = o.write_pages()
        assert G.digest(pages) == want[name], name
        tids, cnt = fixture_tids(o)
        mine = P.write_pages(P.OPCLASSES[opclass][0], dim, m, efc, o.export_graph(), tids, cnt)
        assert hashlib.sha256(mine.tobytes()).hexdigest() == want[name]["sha256"], name
    # a vacuumed relation has no flush-layout twin (deleted values are zeroed in place): the oracle's bytes only, and
    # what the exporter reads out of them
    ov = G.vacuumed_relation()
    pv = ov.write_pages()
    assert G.digest(pv) == want["vector_l2_vacuumed"]
    g = P.export_pages(pv, P.ELEM_F32)
    assert np.array_equal(g["flags"], ov.element_flags()) and g["n_deleted"] == int((ov.element_flags() & 1).sum())
