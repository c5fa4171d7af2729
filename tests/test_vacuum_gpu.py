"""GPU parity of scans over an index a vacuum has touched: elements flagged deleted are
visited but never evaluated or returned (scan.rs:179), a deleted entry point returns no rows (scan.rs:475-486),
elements left without heap TIDs are searched through but emit nothing and do not count towards
hnsw.max_scan_tuples (scan.rs:866-871), stale neighbour tuples read as empty lists (scan.rs:262).  The GPU index is
loaded from the relation's PAGES through the product's exporter; the oracle holds the same states in memory."""
import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import assert_same_results, gpu_index, make_rows, vacuumed_oracle

pytestmark = pytest.mark.gpu


def load_from_pages(o, opclass, dim, m, efc):
    g = P.export_pages(o.write_pages(), P.OPCLASSES[opclass][0])
    assert np.array_equal(g["flags"], o.element_flags())
    ix = gpu_index(opclass, dim, g, m=m, efc=efc)
    ix.set_element_flags(g["flags"])
    return ix, g


@pytest.mark.parametrize("opclass,kind,dim,n", [("vector_l2_ops", "f32", 32, 4000), ("halfvec_cosine_ops", "f16", 48, 1500),
                                                ("bit_hamming_ops", "bit", 256, 2000)])
def test_one_shot_scan(opclass, kind, dim, n):
    m, efc = 8, 32
    o, deleted, no_tid, stale = vacuumed_oracle(opclass, kind, n, dim, m, efc)
    ix, g = load_from_pages(o, opclass, dim, m, efc)
    q = make_rows(kind, 300, dim, 77)
    qn = ix.prepare_queries(q)
    for ef, k in ((40, 10), (10, 10), (100, 60), (1, 1)):
        got = ix.search(q, k=k, ef_search=ef, with_stats=True)
        want = o.search_batch(qn, ef_search=ef, k=k, with_counters=True)
        assert_same_results(got, want, "%s ef=%d" % (opclass, ef))
        for key in ("n_dist", "n_expand0", "n_expand_upper", "n_visited"):
            assert got[3][key] == want[3][key], (key, got[3][key], want[3][key])
        ids = got[0][got[0] >= 0]
        assert not np.isin(ids, deleted).any() and not np.isin(ids, no_tid).any()
    # without the flags the same graph returns flagged elements: the flags are what removes them
    ix.set_element_flags(None)
    ids = ix.search(q, k=10, ef_search=40)[0]
    assert np.isin(ids[ids >= 0], no_tid).any()
    ix.close()


def test_deleted_entry_point_returns_nothing():
    o, *_ = vacuumed_oracle("vector_l2_ops", "f32", 800, 16, 8, 32, delete_entry=True)
    ix, g = load_from_pages(o, "vector_l2_ops", 16, 8, 32)
    q = make_rows("f32", 50, 16, 5)
    ids, dists, counts = ix.search(q, k=10, ef_search=40)
    assert (counts == 0).all() and (ids == -1).all() and np.isinf(dists).all()
    sc = ix.scan_begin(q, ef_search=40, iterative_scan=P.ITER_RELAXED)
    ids, dists, counts, done = sc.next()
    assert (counts == 0).all() and done.all()
    sc.close()
    ix.close()


@pytest.mark.parametrize("mode", [P.ITER_OFF, P.ITER_RELAXED, P.ITER_STRICT])
@pytest.mark.parametrize("max_tuples", [20000, 150])
def test_iterative_stream(mode, max_tuples):
    opclass, kind, dim, m, efc = "vector_l2_ops", "f32", 24, 8, 32
    o, deleted, no_tid, stale = vacuumed_oracle(opclass, kind, 3000, dim, m, efc, frac=0.08)
    ix, g = load_from_pages(o, opclass, dim, m, efc)
    q = make_rows(kind, 16, dim, 78)
    sc = ix.scan_begin(q, ef_search=40, iterative_scan=mode, max_scan_tuples=max_tuples)
    got_i, got_d = sc.stream(None)
    tot = dict(n_dist=0, n_expand0=0, n_expand_upper=0, n_visited=0)
    for i in range(len(q)):
        tids, eids, ds, c = o.scan_stream(q[i], ef_search=40, iterative=mode, max_scan_tuples=max_tuples)
        assert len(got_i[i]) == len(eids), (i, len(got_i[i]), len(eids))
        assert np.array_equal(got_i[i], eids) and np.array_equal(got_d[i], ds.astype(np.float32))
        assert not np.isin(eids, deleted).any() and not np.isin(eids, no_tid).any()
        for key in tot:
            tot[key] += c[key]
    for key in tot:
        assert sc.totals[key] == tot[key], (key, sc.totals[key], tot[key])
    sc.close()
    ix.close()


def test_flags_api_errors_and_build_refusal():
    o, *_ = vacuumed_oracle("vector_l2_ops", "f32", 300, 8, 4, 16)
    ix, g = load_from_pages(o, "vector_l2_ops", 8, 4, 16)
    with pytest.raises(P.PgvrxError) as e:
        ix.set_element_flags(np.zeros(5, np.uint8))
    assert e.value.code == P.ERR_INVALID
    with pytest.raises(P.PgvrxError) as e:
        ix.set_element_flags(np.full(o.n, 8, np.uint8))
    assert e.value.code == P.ERR_INVALID
    ix.close()
    # a handle built in memory refuses to grow once it carries vacuum flags
    rows = make_rows("f32", 200, 8, 3)
    b = P.HnswIndex("vector_l2_ops", 8, m=4, ef_construction=16, capacity=400)
    b.build(rows[:100], np.zeros(100, np.int32))
    f = np.zeros(100, np.uint8)
    f[7] = P.ELEM_NO_HEAPTIDS
    b.set_element_flags(f)
    with pytest.raises(P.PgvrxError) as e:
        b.build(rows[100:], np.zeros(100, np.int32))
    assert e.value.code == P.ERR_STATE
    b.set_element_flags(None)
    b.build(rows[100:], np.zeros(100, np.int32))
    b.close()


@pytest.mark.parametrize("opclass,kind,dim", [("vector_l2_ops", "f32", 16), ("bit_hamming_ops", "bit", 64)])
def test_scan_after_a_real_vacuum(opclass, kind, dim):
    # the oracle's restatement of ambulkdelete (src/index/vacuum.rs: remove dead heap TIDs, repair with skip-set searches,
    # mark deleted) produces the relation; the GPU loads its pages and must scan it like the oracle does
    from helpers import oracle_index
    m, efc, n = 6, 24, 3000
    o = oracle_index(opclass, dim, make_rows(kind, n, dim, 41), m=m, efc=efc)
    rng = np.random.default_rng(42)
    dead = rng.permutation(n)[: n // 3]
    assert o.bulk_delete(dead) == len(dead)
    ix, g = load_from_pages(o, opclass, dim, m, efc)
    assert g["n_deleted"] > 0 and (g["flags"] != 0).sum() == g["n_deleted"]
    q = make_rows(kind, 200, dim, 43)
    qn = ix.prepare_queries(q)
    got = ix.search(q, k=10, ef_search=40, with_stats=True)
    want = o.search_batch(qn, ef_search=40, k=10, with_counters=True)
    assert_same_results(got, want, opclass + " after vacuum")
    for key in ("n_dist", "n_expand0", "n_expand_upper", "n_visited"):
        assert got[3][key] == want[3][key], key
    assert not np.isin(got[0][got[0] >= 0], np.nonzero(g["flags"])[0]).any()
    sc = ix.scan_begin(q[:8], ef_search=20, iterative_scan=P.ITER_RELAXED, max_scan_tuples=300)
    got_i, got_d = sc.stream(None)
    sc.close()
    for i in range(8):
        _, eids, ds, _ = o.scan_stream(qn[i], ef_search=20, iterative=O.ITER_RELAXED, max_scan_tuples=300)
        assert np.array_equal(got_i[i], eids) and np.array_equal(got_d[i], ds.astype(np.float32))
    ix.close()
