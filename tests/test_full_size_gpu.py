"""BASELINE configs[1] at full size (1M x 128 fp32, m=16, ef_construction=64) through size-independent properties:
results sorted by (distance, id), ids distinct and in range, exactly k of them, repeatable; scan == oracle on a sample;
exact kNN is a lower bound of every scan distance, rank by rank; flat results agree between the two flat kernels.
Needs the cached reference-exact graph (data_cache/c2_1m, built by the oracle; bench.py builds and checks the same
graph on the GPU)."""
import json
import os

import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from pgvector_rx_b200 import datasets as D

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = os.path.join(ROOT, "data_cache", "c2_1m")


@pytest.fixture(scope="module")
def c2():
    if not os.path.exists(os.path.join(CACHE, "meta.json")):
        pytest.skip("data_cache/c2_1m not present")
    meta = json.load(open(os.path.join(CACHE, "meta.json")))
    g = {k: np.load(os.path.join(CACHE, k + ".npy")) for k in ("level", "adj0", "upper_base", "adj_up")}
    rows = D.gaussian_f32(meta["n"], meta["dim"], meta["data_seed"])
    ix = P.HnswIndex("vector_l2_ops", meta["dim"], m=meta["m"], ef_construction=meta["efc"], capacity=meta["n"])
    ix.load_graph(rows, g["level"], g["adj0"], g["upper_base"], g["adj_up"], meta["entry"])
    g.update(rows=rows, entry=meta["entry"], n=meta["n"])
    yield ix, g, meta
    ix.close()


def test_scan_properties_at_1m(c2):
    ix, g, meta = c2
    q = D.gaussian_f32(4000, meta["dim"], 777)
    for ef, k in ((40, 10), (100, 100)):
        ids, dists, counts, st = ix.search(q, k=k, ef_search=ef, with_stats=True)
        assert (counts == k).all() and st["n_retried"] == 0
        assert ((ids >= 0) & (ids < meta["n"])).all()
        assert all(len(set(r.tolist())) == k for r in ids[:500])
        assert (np.diff(dists, axis=1) >= 0).all()                                 # nearest first
        ties = np.diff(dists, axis=1) == 0
        assert (np.diff(ids, axis=1)[ties] > 0).all()                              # equal distances: ascending id
        ids2, dists2, _ = ix.search(q, k=k, ef_search=ef)
        assert np.array_equal(ids, ids2) and np.array_equal(dists, dists2)        # idempotent
        # the reported distance is the reference's arithmetic on the reported row
        for i in range(0, 4000, 997):
            for j in (0, k - 1):
                assert np.float32(O.distance(O.ELEM_F32, O.L2SQ, meta["dim"], q[i], g["rows"][ids[i, j]])) == dists[i, j]


def test_scan_equals_oracle_sample_at_1m(c2):
    ix, g, meta = c2
    q = D.gaussian_f32(300, meta["dim"], 778)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, meta["dim"], m=meta["m"], ef_construction=meta["efc"], tie_mode=O.TIE_TOTAL_ORDER)
    o.import_graph(g)
    for ef in (40, 100):
        got = ix.search(q, k=10, ef_search=ef, with_stats=True)
        want = o.search_batch(q, ef_search=ef, k=10, nthreads=os.cpu_count() or 1, with_counters=True)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        for key in ("n_dist", "n_expand0", "n_expand_upper", "n_visited"):
            assert got[3][key] == want[3][key]


def test_flat_bounds_scan_and_paths_agree_at_1m(c2):
    ix, g, meta = c2
    q = D.gaussian_f32(2048, meta["dim"], 779)
    fi, fd, st = ix.flat_topk(q, k=10, with_stats=True)
    assert st["path"] in (1, 2) and st["n_uncertified"] <= 20
    ids, dists, _ = ix.search(q, k=10, ef_search=40)
    assert (fd <= dists).all()                               # the exact j-th neighbour is never further than the scan's j-th
    old = os.environ.get("PGVRX_FLAT_PATH")
    os.environ["PGVRX_FLAT_PATH"] = "exact"
    try:
        ei, ed = ix.flat_topk(q[:256], k=10)
    finally:
        if old is None:
            del os.environ["PGVRX_FLAT_PATH"]
        else:
            os.environ["PGVRX_FLAT_PATH"] = old
    assert np.array_equal(ei, fi[:256]) and np.array_equal(ed, fd[:256])
    wi, wd = O.flat_topk(O.ELEM_F32, O.L2SQ, meta["dim"], g["rows"], q[:24], 10, nthreads=os.cpu_count() or 1)
    assert np.array_equal(wi, fi[:24]) and np.array_equal(wd, fd[:24])
