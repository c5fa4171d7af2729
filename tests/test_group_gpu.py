"""pgvrx_group_*: several GPUs of one box behind one handle.  On a one-GPU box the members share the device (an ordinal
may repeat), which exercises the whole code path -- shard bounds, global ids, the merge over per-member lists, query
slices written into one caller buffer; with two or more GPUs the merge reads the peers' lists over NVLink."""
import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import assert_same_results, make_rows, oracle_index

pytestmark = pytest.mark.gpu


def devices(n):
    nd = P.lib().pgvrx_device_count()
    return [i % nd for i in range(n)]


@pytest.mark.parametrize("opclass,kind,dim,n,nmem", [("vector_l2_ops", "f32", 32, 30000, 3), ("vector_ip_ops", "f32", 24, 20011, 4),
                                                     ("bit_hamming_ops", "bit", 256, 9000, 2), ("halfvec_l2_ops", "f16", 40, 5000, 5)])
def test_sharded_flat_equals_the_whole_table(opclass, kind, dim, n, nmem):
    elem, metric, _ = P.OPCLASSES[opclass]
    rows = make_rows(kind, n, dim, 61)
    q = make_rows(kind, 700, dim, 62)
    g = P.HnswGroup(opclass, dim, devices(nmem), capacity=n)
    g.load_rows(rows)
    gi, gd, st = g.flat_topk(q, k=10, with_stats=True)
    assert st["n_devices"] == nmem and st["n_peer_direct"] >= 1
    one = P.HnswIndex(opclass, dim, capacity=n)
    one.load_rows_stored(rows)
    wi, wd = one.flat_topk(q, k=10, prepared=True)
    assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
    oi, od = O.flat_topk(elem, metric, dim, rows, q[:40], 10)
    assert np.array_equal(gi[:40], oi) and np.array_equal(gd[:40], od)
    one.close()
    g.close()


def test_sharded_flat_more_members_than_rows_and_k_above_shard_size():
    rows = make_rows("f32", 7, 8, 63)
    q = make_rows("f32", 5, 8, 64)
    g = P.HnswGroup("vector_l2_ops", 8, devices(3), capacity=16)
    g.load_rows(rows)
    gi, gd = g.flat_topk(q, k=10)
    oi, od = O.flat_topk(O.ELEM_F32, O.L2SQ, 8, rows, q, 7)
    assert np.array_equal(gi[:, :7], oi) and np.array_equal(gd[:, :7], od) and (gi[:, 7:] == -1).all() and np.isinf(gd[:, 7:]).all()
    g.close()


def test_query_sharded_scan_fills_one_caller_buffer():
    n, dim = 6000, 48
    o = oracle_index("vector_l2_ops", dim, make_rows("f32", n, dim, 65))
    og = o.export_graph()
    q = make_rows("f32", 1001, dim, 66)                      # not a multiple of the member count
    g = P.HnswGroup("vector_l2_ops", dim, devices(3), capacity=n)
    g.load_graph(og["rows"], og["level"], og["adj0"], og["upper_base"], og["adj_up"], og["entry"])
    got = g.search(q, k=10, ef_search=40, with_stats=True)
    assert_same_results(got, o.search_batch(q, ef_search=40, k=10), "group scan")
    assert got[3]["n_devices"] == 3
    with pytest.raises(P.PgvrxError):
        g.flat_topk(q, k=10)                                 # the table is replicated, not sharded
    g.close()


def test_group_on_distinct_devices_reads_peers_in_place():
    if P.lib().pgvrx_device_count() < 2:
        pytest.skip("one GPU")
    nd = min(P.lib().pgvrx_device_count(), 8)
    n, dim = 40000, 64
    rows = make_rows("f32", n, dim, 67)
    q = make_rows("f32", 600, dim, 68)
    g = P.HnswGroup("vector_l2_ops", dim, list(range(nd)), capacity=n)
    g.load_rows(rows)
    gi, gd, st = g.flat_topk(q, k=10, with_stats=True)
    oi, od = O.flat_topk(O.ELEM_F32, O.L2SQ, dim, rows, q[:64], 10)
    assert np.array_equal(gi[:64], oi) and np.array_equal(gd[:64], od)
    assert st["n_peer_direct"] == nd, g.peer_access()
    g.close()
