"""The narrow build seam: pgvrx_search_layer_batch + pgvrx_commit_insert, for a caller that keeps
select_neighbors / update_neighbor_connections on its side (src/graph/mod.rs:144-145,386,407; build.rs:464,514).

* search_layer for a batch must equal the oracle's search_layer restatement (graph/mod.rs:161-255) on the same graph,
  entry points, ef and layer -- ids and distance bits;
* a build driven from the HOST side through the seam -- the oracle plays the caller's heuristic and commit, element by
  element -- leaves exactly the oracle's graph on the device, and scans of it are identical."""
import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import assert_same_results, make_rows, oracle_index
from pgvector_rx_b200 import datasets as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("opclass,kind,dim", [("vector_l2_ops", "f32", 24), ("halfvec_ip_ops", "f16", 40), ("bit_hamming_ops", "bit", 192)])
def test_search_layer_batch_equals_the_oracle(opclass, kind, dim):
    elem, metric, _ = P.OPCLASSES[opclass]
    n, m = 3000, 6
    rows = make_rows(kind, n, dim, 71)
    o = oracle_index(opclass, dim, rows, m=m, efc=24)
    g = o.export_graph()
    ix = P.HnswIndex(opclass, dim, m=m, ef_construction=24, capacity=n)
    ix.load_graph(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    rng = np.random.default_rng(72)
    q = make_rows(kind, 48, dim, 73)
    for layer, ef, n_ep in ((0, 24, 1), (0, 40, 7), (1, 1, 1), (1, 12, 4), (0, 300, 3)):
        on_layer = np.nonzero(g["level"] >= layer)[0]
        pitch = 8
        ep_ids = np.full((len(q), pitch), -1, np.int32)
        ep_d = np.zeros((len(q), pitch), np.float32)
        cnt = np.full(len(q), n_ep, np.int32)
        for i in range(len(q)):
            ids = rng.choice(on_layer, size=n_ep, replace=False)
            ep_ids[i, :n_ep] = ids
            ep_d[i, :n_ep] = [np.float32(O.distance(elem, metric, dim, q[i], g["rows"][j])) for j in ids]
        gi, gd, gc, ex, nex = ix.search_layer_batch(ep_ids, ep_d, cnt, ef, layer, queries=q, exp_cap=64)
        for i in range(len(q)):
            wi, wd = o.search_layer(q[i], ep_ids[i, :n_ep], ep_d[i, :n_ep], ef, layer)
            assert gc[i] == len(wi), (layer, ef, i)
            assert np.array_equal(gi[i, :gc[i]], wi) and np.array_equal(gd[i, :gc[i]], wd), (layer, ef, i)
            assert (gi[i, gc[i]:] == -1).all()
            # the read set: expansion order starts at the nearest entry point (ties: lowest id), every id is on the layer
            k = min(nex[i], 64)
            assert nex[i] >= 1 and (g["level"][ex[i, :k]] >= layer).all()
            first = min(range(n_ep), key=lambda t: (ep_d[i, t], ep_ids[i, t]))
            assert ex[i, 0] == ep_ids[i, first]
    # queries given as element ids: the stored row is the query (the build's case)
    qid = rng.choice(n, size=16, replace=False).astype(np.int32)
    ep_ids = np.full((16, 1), g["entry"], np.int32)
    ep_d = np.array([[np.float32(O.distance(elem, metric, dim, g["rows"][j], g["rows"][g["entry"]]))] for j in qid], np.float32)
    gi, gd, gc = ix.search_layer_batch(ep_ids, ep_d, np.ones(16, np.int32), 24, 0, query_ids=qid)
    for i in range(16):
        wi, wd = o.search_layer(g["rows"][qid[i]], ep_ids[i], ep_d[i], 24, 0)
        assert np.array_equal(gi[i, :gc[i]], wi) and np.array_equal(gd[i, :gc[i]], wd)
    with pytest.raises(P.PgvrxError):
        ix.search_layer_batch(np.array([[int(np.nonzero(g["level"] == 0)[0][0])]], np.int32), np.zeros((1, 1), np.float32),
                              np.ones(1, np.int32), 4, 1, queries=q[:1])          # entry point below the searched layer
    ix.close()


def lists_of(g, e, m):
    lv = int(g["level"][e])
    out = [g["adj0"][e]]
    for lc in range(1, lv + 1):
        out.append(g["adj_up"][g["upper_base"][e] + lc - 1])
    return np.concatenate(out).astype(np.int32)


def test_host_driven_build_through_the_seam():
    n0, n, dim, m, efc = 400, 520, 16, 5, 20
    rows = make_rows("f32", n, dim, 75)
    lv = D.hnsw_levels(n, m, 76)
    lv[n0 + 30] = int(lv[:n0].max()) + 1                      # one insert raises the entry point
    o = O.Oracle(O.ELEM_F32, O.L2SQ, dim, m=m, ef_construction=efc)
    o.build(rows[:n0], lv[:n0])
    g = o.export_graph()
    ix = P.HnswIndex("vector_l2_ops", dim, m=m, ef_construction=efc, capacity=n)
    ix.load_graph(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    for e in range(n0, n):
        # the caller's side: the searches of find_element_neighbors on the GPU (layer 0 shown: entry points = the result
        # of the layer above, here the oracle's own descent), then heuristic + commit -- played by the oracle
        entry = int(g["entry"])
        d0 = np.float32(O.distance(O.ELEM_F32, O.L2SQ, dim, rows[e], g["rows"][entry]))
        top = int(g["level"][entry])
        ep_i, ep_d = np.array([entry], np.int32), np.array([d0], np.float32)
        for lc in range(top, 0, -1):                          # greedy descent (ef = 1 above the new element's level, efc on its layers)
            ef = 1 if lc > lv[e] else efc
            wi, wd = o.search_layer(rows[e], ep_i, ep_d, ef, lc)
            gi, gd, gc = ix.search_layer_batch(ep_i[None, :], ep_d[None, :], np.array([len(ep_i)], np.int32), ef, lc, queries=rows[e:e + 1])
            assert np.array_equal(gi[0, :gc[0]], wi) and np.array_equal(gd[0, :gc[0]], wd)
            ep_i, ep_d = (wi[:1], wd[:1]) if lc > lv[e] else (wi, wd)
        wi, wd = o.search_layer(rows[e], ep_i, ep_d, efc, 0)
        gi, gd, gc = ix.search_layer_batch(ep_i[None, :], ep_d[None, :], np.array([len(ep_i)], np.int32), efc, 0, queries=rows[e:e + 1])
        assert np.array_equal(gi[0, :gc[0]], wi) and np.array_equal(gd[0, :gc[0]], wd)
        o.build(rows[e:e + 1], lv[e:e + 1], base_tid=e)
        g2 = o.export_graph()
        assert g2["n"] == e + 1
        mod_nodes, mod_layers, mod_lists = [], [], []
        for v in np.nonzero((g2["adj0"][:e] != g["adj0"]).any(axis=1))[0]:
            mod_nodes.append(v); mod_layers.append(0); mod_lists.append(g2["adj0"][v])
        for v in np.nonzero(g["level"] > 0)[0]:
            for lc in range(1, int(g["level"][v]) + 1):
                a, b = g["adj_up"][g["upper_base"][v] + lc - 1], g2["adj_up"][g2["upper_base"][v] + lc - 1]
                if not np.array_equal(a, b):
                    mod_nodes.append(v); mod_layers.append(lc); mod_lists.append(b)
        new_id = ix.commit_insert(rows[e], int(lv[e]), lists_of(g2, e, m), mod_nodes, mod_layers,
                                  np.concatenate(mod_lists) if mod_lists else np.zeros(0, np.int32))
        assert new_id == e
        g = g2
    got = ix.export_graph()
    for key in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(got[key], g[key]), key
    assert got["entry"] == g["entry"] == n0 + 30
    q = make_rows("f32", 200, dim, 77)
    assert_same_results(ix.search(q, k=10, ef_search=40), o.search_batch(q, ef_search=40, k=10), "scan after a host-driven build")
    # the first element of an empty index through the same call
    ex = P.HnswIndex("vector_l2_ops", dim, m=m, ef_construction=efc, capacity=4)
    assert ex.commit_insert(rows[0], 2, np.full(2 * m + 2 * m, -1, np.int32)) == 0
    assert ex.graph_info()["entry"] == 0 and ex.graph_info()["entry_level"] == 2
    ex.close()
    ix.close()
