"""GPU parity of inserts into a live index (aminsert, src/index/insert.rs:1227-1480): the same
rows inserted through pgvrx_index_insert and through the oracle's restatement of aminsert must leave IDENTICAL graphs --
every ordered neighbour list of every layer, the entry point and the heap TIDs -- although the GPU path searches whole
batches speculatively and commits them through the replay rule.  The rules differ from CREATE INDEX: nearest-lm
lists for the new element, one-slot replacement in a full neighbour list (get_update_index, insert.rs:500-738)."""
import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import assert_same_results, gpu_index, make_rows, oracle_index
from pgvector_rx_b200 import datasets as D

pytestmark = pytest.mark.gpu


def assert_same_graph(ix, o, what):
    g, w = ix.export_graph(), o.export_graph()
    assert g["n"] == w["n"] and g["entry"] == w["entry"], (what, g["n"], w["n"], g["entry"], w["entry"])
    assert np.array_equal(g["level"], w["level"]) and np.array_equal(g["upper_base"], w["upper_base"])
    bad = np.nonzero((g["adj0"] != w["adj0"]).any(axis=1))[0]
    assert len(bad) == 0, "%s: %d layer-0 lists differ, first %d:\n gpu %s\n orc %s" % (
        what, len(bad), bad[0], g["adj0"][bad[0]], w["adj0"][bad[0]])
    assert np.array_equal(g["adj_up"], w["adj_up"]), what
    assert np.array_equal(g["rows"].view(np.uint8), w["rows"].view(np.uint8))
    return g


def run_case(opclass, kind, n0, n1, dim, m=8, efc=32, seed=31, base="oracle", chunks=1):
    """n0 rows indexed by CREATE INDEX (oracle-built and loaded, or built on the GPU), then n1 rows inserted."""
    elem, metric, norm = P.OPCLASSES[opclass]
    rows = make_rows(kind, n0 + n1, dim, seed)
    lv = D.hnsw_levels(n0 + n1, m, seed + 1)
    o = O.Oracle(elem, metric, dim, m=m, ef_construction=efc, tie_mode=O.TIE_TOTAL_ORDER)
    ix = P.HnswIndex(opclass, dim, m=m, ef_construction=efc, capacity=n0 + n1 + 1)
    r0 = rows[:n0]
    if norm:
        r0 = P.normalize_rows(elem, dim, r0)[1]
        rows = np.concatenate([r0, P.normalize_rows(elem, dim, rows[n0:])[1]])
    if n0:
        o.build(rows[:n0], lv[:n0])
        if base == "oracle":
            g0 = o.export_graph()
            ix.load_graph(g0["rows"], g0["level"], g0["adj0"], g0["upper_base"], g0["adj_up"], g0["entry"])
        else:
            assert not norm               # build() would renormalise rows that are already unit length
            ix.build(rows[:n0], lv[:n0])
    step = (n1 + chunks - 1) // chunks
    stats = dict(n_rounds=0, n_conflicts=0, n_host_dist=0, n_searched=0)
    for c0 in range(n0, n0 + n1, step):
        c1 = min(n0 + n1, c0 + step)
        o.insert(rows[c0:c1], lv[c0:c1], base_tid=c0)
        st = _insert_raw(ix, rows[c0:c1], lv[c0:c1], np.arange(c0, c1, dtype=np.uint64))
        for k in stats:
            stats[k] += st[k]
    g = assert_same_graph(ix, o, "%s n0=%d n1=%d" % (opclass, n0, n1))
    for i in range(0, o.n, max(1, o.n // 50)):
        assert g["heap_tid_counts"][i] == len(o.heap_tids(i))
    # and the scan over the grown index matches too
    q = make_rows(kind, 100, dim, seed + 2)
    qn = ix.prepare_queries(q)
    assert_same_results(ix.search(q, k=10, ef_search=40), o.search_batch(qn, ef_search=40, k=10), opclass)
    ix.close()
    return stats


def _insert_raw(ix, rows, levels, tids):
    """pgvrx_index_insert without the binding's normalisation (rows are already prepared)."""
    import ctypes as C
    rows = P._rows(ix.elem, ix.dim, rows)
    levels = np.ascontiguousarray(levels, np.int32)
    st = P.BuildStats()
    ix._check(ix.L.pgvrx_index_insert(ix.h, P._p(rows), len(rows), P._p(levels), P._p(tids), 0, C.byref(st)))
    return st.as_dict()


def test_insert_into_oracle_built_index_l2():
    st = run_case("vector_l2_ops", "f32", 3000, 2000, 32)
    assert st["n_rounds"] < 2000          # batches, not one insert per round
    assert st["n_host_dist"] == 0         # the pairwise tests ran on the GPU


def test_insert_into_gpu_built_index_and_in_chunks():
    run_case("vector_l2_ops", "f32", 1500, 1500, 24, base="gpu", chunks=3)


def test_insert_into_empty_index():
    run_case("vector_l2_ops", "f32", 0, 2500, 16, m=6, efc=24)


@pytest.mark.parametrize("opclass,kind,dim", [("vector_ip_ops", "f32", 48), ("vector_cosine_ops", "f32", 20),
                                              ("halfvec_l2_ops", "f16", 40), ("vector_l1_ops", "f32", 12),
                                              ("bit_hamming_ops", "bit", 128), ("bit_jaccard_ops", "bit", 200)])
def test_insert_other_opclasses(opclass, kind, dim):
    run_case(opclass, kind, 800, 1200, dim)


@pytest.mark.parametrize("m,efc", [(2, 4), (16, 64), (24, 100)])
def test_insert_m_range(m, efc):
    run_case("vector_l2_ops", "f32", 500, 1500, 16, m=m, efc=efc)


def test_insert_duplicates_and_level_changes():
    # duplicates merge into an element with room (insert.rs:1136-1213); a higher level moves the entry point
    rng = np.random.default_rng(3)
    base = rng.standard_normal((300, 8)).astype(np.float32)
    rows = np.concatenate([base, base[:80], base[:80], np.tile(base[5], (14, 1)), rng.standard_normal((200, 8)).astype(np.float32)])
    lv = np.zeros(len(rows), np.int32)
    lv[::23] = 1
    lv[350] = 3
    lv[600] = 5
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 8, m=4, ef_construction=16, tie_mode=O.TIE_TOTAL_ORDER)
    ix = P.HnswIndex("vector_l2_ops", 8, m=4, ef_construction=16, capacity=len(rows))
    o.build(rows[:200], lv[:200])
    ix.build(rows[:200], lv[:200])
    o.insert(rows[200:], lv[200:], base_tid=200)
    st = _insert_raw(ix, rows[200:], lv[200:], np.arange(200, len(rows), dtype=np.uint64))
    g = assert_same_graph(ix, o, "duplicates")
    assert st["n_duplicates"] > 150
    for i in range(o.n):
        t = o.heap_tids(i)
        assert g["heap_tid_counts"][i] == len(t) and g["heap_tids"][i, :len(t)].astype(np.int64).tolist() == t.tolist()
    assert g["heap_tid_counts"].max() == 10
    ix.close()


def _same_live_state(ix, o, what):
    g, w = ix.export_graph(), o.export_graph()
    assert g["n"] == w["n"] and g["entry"] == w["entry"], (what, g["n"], w["n"], g["entry"], w["entry"])
    bad = np.nonzero((g["adj0"] != w["adj0"]).any(axis=1))[0]
    assert len(bad) == 0, "%s: %d layer-0 lists differ, first %d:\n gpu %s\n orc %s" % (
        what, len(bad), bad[0], g["adj0"][bad[0]], w["adj0"][bad[0]])
    assert np.array_equal(g["adj_up"], w["adj_up"]) and np.array_equal(g["level"], w["level"]), what
    assert np.array_equal(ix.element_flags(), o.element_flags()), what
    assert np.array_equal(g["rows"].view(np.uint8), w["rows"].view(np.uint8)), what
    for i in range(o.n):
        t = o.heap_tids(i)
        assert g["heap_tid_counts"][i] == len(t) and g["heap_tids"][i, :len(t)].astype(np.int64).tolist() == t.tolist(), (what, i)


def _life_cycle(opclass, kind, n0, dim, m, efc, seed, kill_entry=False, dup_rows=0):
    """CREATE INDEX -> VACUUM -> INSERT -> scan -> VACUUM -> INSERT -> scan, every step against the oracle: the normal
    life of an index (aminsert on a relation that holds deleted elements: insert.rs:525-528, 569-623, 1036-1047)."""
    elem, metric, _ = P.OPCLASSES[opclass]
    n_new = n0 // 2
    rows = make_rows(kind, n0 + 2 * n_new, dim, seed)
    if dup_rows:
        rows[n0:n0 + dup_rows] = rows[:dup_rows]                 # re-inserted values: some of their twins are dead by then
    lv = D.hnsw_levels(len(rows), m, seed + 1)
    o = O.Oracle(elem, metric, dim, m=m, ef_construction=efc, tie_mode=O.TIE_TOTAL_ORDER)
    ix = P.HnswIndex(opclass, dim, m=m, ef_construction=efc, capacity=len(rows))
    o.build(rows[:n0], lv[:n0])
    ix.build(rows[:n0], lv[:n0], heap_tids=np.arange(n0, dtype=np.uint64))
    rng = np.random.default_rng(seed)
    q = make_rows(kind, 64, dim, seed + 2)
    alive = np.arange(n0)
    lo = n0
    for cycle in range(2):
        dead = rng.permutation(alive)[: len(alive) // 3]
        if kill_entry and cycle == 0 and o.entry not in dead:
            dead = np.append(dead, o.entry)
        alive = np.setdiff1d(alive, dead)
        assert o.bulk_delete(dead) == ix.vacuum(dead)["n_heap_tids_removed"]
        _same_live_state(ix, o, "%s vacuum %d" % (opclass, cycle))
        hi = lo + n_new
        o.insert(rows[lo:hi], lv[lo:hi], base_tid=lo)
        st = _insert_raw(ix, rows[lo:hi], lv[lo:hi], np.arange(lo, hi, dtype=np.uint64))
        assert st["n_host_dist"] == 0
        _same_live_state(ix, o, "%s insert after vacuum %d" % (opclass, cycle))
        alive = np.concatenate([alive, np.arange(lo, hi)])       # (heap TID t <-> input row t; merged duplicates keep their TID)
        assert_same_results(ix.search(q, k=10, ef_search=40), o.search_batch(q, ef_search=40, k=10), "scan, cycle %d" % cycle)
        lo = hi
    ix.close()


def test_insert_into_vacuumed_index_l2():
    _life_cycle("vector_l2_ops", "f32", 1500, 12, 6, 24, seed=71)


def test_insert_into_vacuumed_index_entry_point_deleted_and_duplicates():
    _life_cycle("vector_l2_ops", "f32", 900, 8, 4, 16, seed=73, kill_entry=True, dup_rows=120)


def test_insert_into_vacuumed_index_bit_hamming():
    _life_cycle("bit_hamming_ops", "bit", 1200, 64, 6, 24, seed=75)


def test_insert_into_index_whose_every_element_was_vacuumed():
    rows = make_rows("f32", 300, 8, 77)
    lv = D.hnsw_levels(300, 4, 78)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 8, m=4, ef_construction=16, tie_mode=O.TIE_TOTAL_ORDER)
    ix = P.HnswIndex("vector_l2_ops", 8, m=4, ef_construction=16, capacity=300)
    o.build(rows[:100], lv[:100])
    ix.build(rows[:100], lv[:100], heap_tids=np.arange(100, dtype=np.uint64))
    dead = np.arange(100)
    assert o.bulk_delete(dead) == ix.vacuum(dead)["n_heap_tids_removed"] == 100
    o.insert(rows[100:], lv[100:], base_tid=100)
    _insert_raw(ix, rows[100:], lv[100:], np.arange(100, 300, dtype=np.uint64))
    _same_live_state(ix, o, "insert into an emptied index")
    q = make_rows("f32", 16, 8, 79)
    assert_same_results(ix.search(q, k=5, ef_search=20), o.search_batch(q, ef_search=20, k=5), "scan of the refilled index")
    ix.close()


def test_insert_into_loaded_index_that_carries_flags():
    # flags that came from a relation's pages (pgvrx_pages_element_flags), not from a vacuum run here: deleted elements are
    # still referenced by live lists, which is where get_update_index's slot reuse applies
    n, dim, m, efc = 800, 10, 4, 16
    rows = make_rows("f32", n + 300, dim, 81)
    lv = D.hnsw_levels(n + 300, m, 82)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, dim, m=m, ef_construction=efc, tie_mode=O.TIE_TOTAL_ORDER)
    o.build(rows[:n], lv[:n])
    rng = np.random.default_rng(83)
    pick = np.setdiff1d(rng.permutation(n)[: n // 4], [o.entry])
    o.vacuum_mark(pick[: len(pick) // 2], O.VACUUM_MARK_DELETED)      # deleted in place: live lists still point at them
    o.vacuum_mark(pick[len(pick) // 2:], O.VACUUM_REMOVE_HEAPTIDS)    # pass 1 of a vacuum only: no heap TID left, not deleted yet
    g = o.export_graph()
    ix = P.HnswIndex("vector_l2_ops", dim, m=m, ef_construction=efc, capacity=n + 300)
    ix.load_graph(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    ix.set_element_flags(o.element_flags())
    o.insert(rows[n:], lv[n:], base_tid=n)
    _insert_raw(ix, rows[n:], lv[n:], np.arange(n, n + 300, dtype=np.uint64))
    gg, w = ix.export_graph(), o.export_graph()
    assert np.array_equal(gg["adj0"], w["adj0"]) and np.array_equal(gg["adj_up"], w["adj_up"]) and gg["entry"] == w["entry"]
    q = make_rows("f32", 32, dim, 84)
    assert_same_results(ix.search(q, k=10, ef_search=40), o.search_batch(q, ef_search=40, k=10), "scan after inserts over flags")
    ix.close()


def test_insert_into_page_exported_index_respects_heap_tid_capacity():
    # an index exported from pages keeps its heap TIDs (pgvrx_index_set_heap_tids): a duplicate row joins an element only
    # while that element holds 1..9 TIDs (insert.rs:1159-1163), exactly as in the oracle that wrote the pages
    rng = np.random.default_rng(11)
    base = rng.standard_normal((150, 8)).astype(np.float32)
    rows0 = np.concatenate([base, np.tile(base[3], (9, 1)), np.tile(base[4], (4, 1))])     # element 3: 10 TIDs, element 4: 5
    lv0 = np.zeros(len(rows0), np.int32)
    lv0[::19] = 1
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 8, m=4, ef_construction=16, tie_mode=O.TIE_TOTAL_ORDER)
    o.build(rows0, lv0)
    assert len(o.heap_tids(3)) == 10 and len(o.heap_tids(4)) == 5
    g = P.export_pages(o.write_pages(), P.ELEM_F32)
    ix = P.HnswIndex("vector_l2_ops", 8, m=4, ef_construction=16, capacity=400)
    ix.load_graph(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    # the fixture writes heap TID t as (block t >> 16, offset (t & 0xffff) + 1); undo the +1 to compare with the oracle
    ix.set_heap_tids(np.where(g["heap_tids"] == np.uint64(0xFFFFFFFFFFFFFFFF), g["heap_tids"], g["heap_tids"] - np.uint64(1)),
                     g["heap_tid_counts"])
    new = np.concatenate([np.tile(base[3], (3, 1)), np.tile(base[4], (7, 1)), rng.standard_normal((40, 8)).astype(np.float32)])
    lv1 = np.zeros(len(new), np.int32)
    t0 = 1000
    o.insert(new, lv1, base_tid=t0)
    _insert_raw(ix, new, lv1, np.arange(t0, t0 + len(new), dtype=np.uint64))
    gg = assert_same_graph(ix, o, "page-exported base")
    for i in range(o.n):
        t = o.heap_tids(i)
        assert gg["heap_tid_counts"][i] == len(t), i
        assert gg["heap_tids"][i, :len(t)].astype(np.int64).tolist() == t.tolist(), i
    ix.close()
