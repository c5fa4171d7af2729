"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py).

CPU part: the oracle still reproduces every frozen output (graph adjacency of a fresh build, scan results and
work counters, exact kNN, iterative streams) -- so the checker itself cannot drift unnoticed.
GPU part: the CUDA path, through the C ABI, reproduces the same frozen outputs bit for bit."""
import os
import sys

import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import make_rows, oracle_index

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import CASES   # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def inputs(name):
    opclass, kind, n, dim, nq, m, efc, ef, k, seed = CASES[name]
    elem, metric, norm = P.OPCLASSES[opclass]
    rows = make_rows(kind, n, dim, seed)
    q = make_rows(kind, nq, dim, seed + 1)
    return rows, q


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    opclass, kind, n, dim, nq, m, efc, ef, k, seed = CASES[name]
    elem, metric, norm = P.OPCLASSES[opclass]
    g = load(name)
    rows, q = inputs(name)
    o = oracle_index(opclass, dim, rows, m=m, efc=efc)
    og = o.export_graph()
    for key in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(og[key], g[key]), key
    assert int(og["entry"]) == int(g["entry"]) and int(og["n"]) == int(g["n"])
    qn = np.stack([O.normalize(elem, dim, r)[1] for r in q]) if norm else q
    ids, dists, counts, ctr = o.search_batch(qn, ef_search=ef, k=k, with_counters=True)
    assert np.array_equal(ids, g["scan_ids"]) and np.array_equal(dists, g["scan_dists"]) and np.array_equal(counts, g["scan_counts"])
    assert [ctr["n_dist"], ctr["n_expand0"], ctr["n_expand_upper"], ctr["n_visited"]] == g["counters"].tolist()
    fi, fd = O.flat_topk(elem, metric, dim, og["rows"], qn, k)
    assert np.array_equal(fi, g["flat_ids"]) and np.array_equal(fd, g["flat_dists"])
    for mode, tag in ((O.ITER_RELAXED, "relaxed"), (O.ITER_STRICT, "strict")):
        for i in range(4):
            _, eids, _, _ = o.scan_stream(qn[i], ef_search=ef, iterative=mode, max_scan_tuples=20000, max_emit=60)
            want = g["iter_" + tag][i]
            assert np.array_equal(eids, want[want >= 0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_golden(name):
    opclass, kind, n, dim, nq, m, efc, ef, k, seed = CASES[name]
    elem, metric, norm = P.OPCLASSES[opclass]
    g = load(name)
    rows, q = inputs(name)
    # build on the GPU: the adjacency must be the frozen one
    ix = P.HnswIndex(opclass, dim, m=m, ef_construction=efc, capacity=len(rows))
    from pgvector_rx_b200 import datasets as D
    ix.build(rows, D.hnsw_levels(len(rows), m, 4))
    bg = ix.export_graph()
    for key in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(bg[key], g[key]), key
    assert int(bg["entry"]) == int(g["entry"]) and int(bg["n"]) == int(g["n"])
    ids, dists, counts, st = ix.search(q, k=k, ef_search=ef, with_stats=True)
    assert np.array_equal(ids, g["scan_ids"]) and np.array_equal(dists, g["scan_dists"]) and np.array_equal(counts, g["scan_counts"])
    assert [st["n_dist"], st["n_expand0"], st["n_expand_upper"], st["n_visited"]] == g["counters"].tolist()
    fi, fd = ix.flat_topk(q, k=k)
    assert np.array_equal(fi, g["flat_ids"]) and np.array_equal(fd, g["flat_dists"])
    for mode, tag in ((P.ITER_RELAXED, "relaxed"), (P.ITER_STRICT, "strict")):
        sc = ix.scan_begin(q[:4], ef_search=ef, iterative_scan=mode)
        got, _ = sc.stream(60)
        for i in range(4):
            want = g["iter_" + tag][i]
            assert np.array_equal(got[i], want[want >= 0])
        sc.close()
    ix.close()


# ---- live index: CREATE INDEX + aminsert, then scans over what a vacuum leaves behind ---------------------------------
from make_golden import LIVE_CASES, live_inputs, live_oracle   # noqa: E402


@pytest.mark.parametrize("name", sorted(LIVE_CASES))
def test_oracle_reproduces_live_golden(name):
    opclass, kind, n0, n1, dim, nq, m, efc, ef, k, seed = LIVE_CASES[name]
    g = load(name)
    o, gi, q = live_oracle(name)
    for key in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(gi[key], g["ins_" + key]), key
    assert int(gi["entry"]) == int(g["ins_entry"]) and int(gi["n"]) == int(g["ins_n"])
    gv = P.export_pages(o.write_pages(), P.OPCLASSES[opclass][0])
    assert np.array_equal(gv["flags"], g["vac_flags"]) and np.array_equal(gv["adj0"], g["vac_adj0"])
    ids, dists, counts, ctr = o.search_batch(q, ef_search=ef, k=k, with_counters=True)
    assert np.array_equal(ids, g["scan_ids"]) and np.array_equal(dists, g["scan_dists"]) and np.array_equal(counts, g["scan_counts"])
    assert [ctr["n_dist"], ctr["n_expand0"], ctr["n_expand_upper"], ctr["n_visited"]] == g["counters"].tolist()
    for mode, tag in ((O.ITER_RELAXED, "relaxed"), (O.ITER_STRICT, "strict")):
        for i in range(4):
            _, eids, _, _ = o.scan_stream(q[i], ef_search=ef, iterative=mode, max_scan_tuples=100, max_emit=150)
            want = g["iter_" + tag][i]
            assert np.array_equal(eids, want[want >= 0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LIVE_CASES))
def test_gpu_reproduces_live_golden(name):
    """No oracle on this path: inputs from seeds, expectations from the committed fixture."""
    opclass, kind, n0, n1, dim, nq, m, efc, ef, k, seed = LIVE_CASES[name]
    g = load(name)
    rows, lv, q = live_inputs(name)
    ix = P.HnswIndex(opclass, dim, m=m, ef_construction=efc, capacity=n0 + n1)
    ix.build(rows[:n0], lv[:n0])
    ix.insert(rows[n0:], lv[n0:], heap_tids=np.arange(n0, n0 + n1, dtype=np.uint64))
    bg = ix.export_graph()
    for key in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(bg[key], g["ins_" + key]), key
    assert int(bg["entry"]) == int(g["ins_entry"]) and int(bg["n"]) == int(g["ins_n"])
    # the vacuumed relation as exported from its pages: same elements, some lists emptied, flags set
    vrows = bg["rows"].copy()
    vrows[(g["vac_flags"] & P.ELEM_DELETED) != 0] = 0
    ix.load_graph(vrows, bg["level"], g["vac_adj0"], bg["upper_base"], g["vac_adj_up"], int(bg["entry"]))
    ix.set_element_flags(g["vac_flags"])
    ids, dists, counts, st = ix.search(q, k=k, ef_search=ef, with_stats=True)
    assert np.array_equal(ids, g["scan_ids"]) and np.array_equal(dists, g["scan_dists"]) and np.array_equal(counts, g["scan_counts"])
    assert [st["n_dist"], st["n_expand0"], st["n_expand_upper"], st["n_visited"]] == g["counters"].tolist()
    for mode, tag in ((P.ITER_RELAXED, "relaxed"), (P.ITER_STRICT, "strict")):
        sc = ix.scan_begin(q[:4], ef_search=ef, iterative_scan=mode, max_scan_tuples=100)
        got, _ = sc.stream(150)
        sc.close()
        for i in range(4):
            want = g["iter_" + tag][i]
            assert np.array_equal(got[i], want[want >= 0]), (tag, i)
    ix.close()
