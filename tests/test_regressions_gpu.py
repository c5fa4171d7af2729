"""Regressions for scratch exhaustion and handle life-cycle rules (GPU, through the C ABI).

* a greedy descent that visits more elements of ONE upper layer than the small per-layer table holds must be answered by
  the retry pass (scan) / the escalated re-run (build), never by an error or an endless loop -- the reference only gets
  slower on such graphs (src/index/scan.rs:492-512, src/graph/mod.rs:385-399);
* a search whose tie tail overflows escalates through larger scratch and finally FAILS with a status instead of repeating
  the round for ever;
* an open iterative scan refuses to continue after the index changed;
* the exact paths skip elements a vacuum has emptied or deleted.
"""
import os

import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import assert_same_results, make_rows, oracle_index
from pgvector_rx_b200 import datasets as D

pytestmark = pytest.mark.gpu


def line_rows(n):
    """points on a line, inserted from one end: every search that starts at element 0 walks the whole line"""
    r = np.zeros((n, 2), np.float32)
    r[:, 0] = np.arange(n, dtype=np.float32)
    return r


def same_graph(a, b):
    return a["n"] == b["n"] and a["entry"] == b["entry"] and all(np.array_equal(a[k], b[k]) for k in ("level", "adj0", "upper_base", "adj_up"))


def test_scan_descent_longer_than_the_small_upper_table():
    n = 3000
    rows = line_rows(n)
    lv = np.ones(n, np.int32)
    lv[0] = 2                                    # the entry point stays at the near end (strictly greater level needed)
    o = oracle_index("vector_l2_ops", 2, rows, levels=lv)
    g = o.export_graph()
    ix = P.HnswIndex("vector_l2_ops", 2, capacity=n)
    ix.load_graph(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    q = np.zeros((64, 2), np.float32)
    q[:, 0] = np.linspace(n - 200, n + 5, 64, dtype=np.float32)      # far end: the layer-1 descent passes > 736 elements
    got = ix.search(q, k=10, ef_search=40, with_stats=True)
    want = o.search_batch(q, ef_search=40, k=10)
    assert_same_results(got, want, "long descent")
    assert got[3]["n_retried"] > 0               # the main pass did overflow; the retry pass answered
    ix.close()


def test_build_descent_longer_than_the_small_upper_table_does_not_hang():
    n, n0 = 2600, 2000
    rows = line_rows(n)
    lv = np.ones(n, np.int32)
    lv[0] = 2
    lv[n0:] = 0                                  # the last rows descend through all of layer 1 with ef = 1
    ix = P.HnswIndex("vector_l2_ops", 2, capacity=n)
    st = ix.build(rows, lv)
    o = oracle_index("vector_l2_ops", 2, rows, levels=lv)
    assert same_graph(ix.export_graph(), o.export_graph())
    assert st["n_big_retries"] > 0
    ix.close()


def test_build_tie_overflow_escalates_and_finally_fails_instead_of_looping():
    # 256 distinct rows, each ~23 times: three elements per value (10 heap TIDs each), C(8,d) * 3 elements at Hamming
    # distance d from any row (84 at d = 2, 168 at d = 3).  Entries evicted from W that still tie with its furthest
    # distance stay expandable (strict `>` stop rule); there can be up to ef - 1 of them, and at ef_construction = 100 the
    # register-resident list keeps 27: dozens of d = 3 entries are evicted one by one while W's furthest stays at 3
    rows = make_rows("bit", 6000, 8, 5)
    lv = D.hnsw_levels(len(rows), 8, 6)
    old = os.environ.get("PGVRX_BUILD_TAILS")
    try:
        os.environ["PGVRX_BUILD_TAILS"] = "30,32,4096"       # levels 1 and 2 overflow again, level 3 holds
        ix = P.HnswIndex("bit_hamming_ops", 8, m=8, ef_construction=100, capacity=len(rows))
        st = ix.build(rows, lv)
        o = oracle_index("bit_hamming_ops", 8, rows, m=8, efc=100, levels=lv)
        assert same_graph(ix.export_graph(), o.export_graph())
        assert st["n_big_retries"] > 0
        ix.close()
        os.environ["PGVRX_BUILD_TAILS"] = "28,29,30"         # nothing holds: the call must fail, not spin
        ix = P.HnswIndex("bit_hamming_ops", 8, m=8, ef_construction=100, capacity=len(rows))
        with pytest.raises(P.PgvrxError) as ei:
            ix.build(rows, lv)
        assert ei.value.code == P.ERR_UNSUPPORTED and "exhausted" in ei.value.message
        # the handle is still consistent: what was committed before the failure is searchable
        n_done = ix.graph_info()["n"]
        assert 0 < n_done < len(rows)
        q = make_rows("bit", 4, 8, 7)
        ids, _, counts = ix.search(q, k=5, ef_search=8)
        assert (counts >= 1).all() and (ids[:, 0] >= 0).all() and (ids < n_done).all()
        ix.close()
    finally:
        if old is None:
            os.environ.pop("PGVRX_BUILD_TAILS", None)
        else:
            os.environ["PGVRX_BUILD_TAILS"] = old


def test_open_scan_is_invalidated_when_the_index_changes():
    n, dim = 1200, 16
    rows = make_rows("f32", n + 300, dim, 41)
    lv = D.hnsw_levels(n + 300, 16, 42)
    ix = P.HnswIndex("vector_l2_ops", dim, capacity=n + 300)
    ix.build(rows[:n], lv[:n])
    q = make_rows("f32", 8, dim, 43)
    sc = ix.scan_begin(q, ef_search=20, iterative_scan=P.ITER_RELAXED)
    sc.next()
    ix.insert(rows[n:], lv[n:], heap_tids=np.arange(n, n + 300, dtype=np.uint64))
    with pytest.raises(P.PgvrxError) as ei:
        sc.next()
    assert ei.value.code == P.ERR_STATE
    sc.close()
    # a scan begun after the change sees the grown index
    sc = ix.scan_begin(q, ef_search=20, iterative_scan=P.ITER_RELAXED)
    ids, _ = sc.stream(50)
    sc.close()
    o = O.Oracle(O.ELEM_F32, O.L2SQ, dim)
    o.build(rows[:n], lv[:n])
    o.insert(rows[n:], lv[n:], base_tid=n)
    for i in range(len(q)):
        _, eids, _, _ = o.scan_stream(q[i], ef_search=20, iterative=O.ITER_RELAXED, max_emit=50)
        assert np.array_equal(ids[i], eids)
    ix.close()


@pytest.mark.parametrize("opclass,kind,dim", [("vector_l2_ops", "f32", 24), ("bit_hamming_ops", "bit", 128)])
def test_exact_paths_skip_elements_a_vacuum_removed(opclass, kind, dim):
    n = 3000
    rows = make_rows(kind, n, dim, 51)
    ix = P.HnswIndex(opclass, dim, capacity=n)
    ix.build(rows, D.hnsw_levels(n, 16, 52))
    dead = np.random.default_rng(53).permutation(n)[:900].astype(np.uint64)
    ix.vacuum(dead)
    flags = ix.element_flags()
    assert (flags != 0).sum() == 900
    q = make_rows(kind, 40, dim, 54)
    gi, gd = ix.flat_topk(q, k=10)
    assert not np.isin(gi, dead.astype(np.int64)).any()
    live = np.nonzero(flags == 0)[0]
    elem, metric, _ = P.OPCLASSES[opclass]
    wi, wd = O.flat_topk(elem, metric, dim, rows[live], q, 10)
    assert np.array_equal(gi, live[wi]) and np.array_equal(gd, wd)
    # re-rank: flagged candidates are dropped
    cand = np.tile(np.arange(64, dtype=np.int32), (len(q), 1))
    ri, _ = ix.rerank(q, cand, k=10)
    assert not np.isin(ri[ri >= 0], dead.astype(np.int64)).any()
    ix.close()
