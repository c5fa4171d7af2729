"""GPU build parity (north star #2): pgvrx_index_build -- GPU-batched search_layer + host heuristic --
must produce EXACTLY the adjacency (ordered lists, every layer), entry point and heap-TID lists of the
oracle's sequential ambuild restatement on the same rows, insert order and levels."""
import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import assert_same_results, make_rows, normalize_rows, oracle_index
from pgvector_rx_b200 import datasets as D

pytestmark = pytest.mark.gpu


def assert_same_graph(got, want, o=None):
    assert got["n"] == want["n"] and got["entry"] == want["entry"]
    for k in ("level", "upper_base", "adj_up", "adj0"):
        if not np.array_equal(got[k], want[k]):
            bad = np.nonzero((got[k] != want[k]).reshape(len(got[k]), -1).any(axis=1))[0]
            raise AssertionError("%s differs at %d rows, first %d:\n gpu %s\n orc %s" % (k, len(bad), bad[0], got[k][bad[0]], want[k][bad[0]]))
    assert np.array_equal(got["rows"].view(np.uint8), want["rows"].view(np.uint8))
    if o is not None:
        for i in range(want["n"]):
            t = o.heap_tids(i)
            assert got["heap_tid_counts"][i] == len(t) and got["heap_tids"][i, :len(t)].tolist() == t.tolist()


def build_both(opclass, kind, n, dim, m=16, efc=64, seed=31, levels=None, rows=None, threads=0):
    elem, metric, norm = P.OPCLASSES[opclass]
    rows = make_rows(kind, n, dim, seed) if rows is None else rows
    lv = D.hnsw_levels(len(rows), m, seed + 1) if levels is None else levels
    ix = P.HnswIndex(opclass, dim, m=m, ef_construction=efc, capacity=len(rows) + 8)
    st = ix.build(rows, lv, host_threads=threads)
    if norm:
        keep = [i for i, r in enumerate(rows) if O.normalize(elem, dim, r)[0] != 0.0]
        orows, olv = normalize_rows(elem, dim, rows), lv[keep]
    else:
        orows, olv = rows, lv
    o = O.Oracle(elem, metric, dim, m=m, ef_construction=efc, tie_mode=O.TIE_TOTAL_ORDER)
    if norm:
        for i, k in enumerate(keep):                      # heap TID = original row number
            o.build(orows[i:i + 1], olv[i:i + 1], base_tid=k)
    else:
        o.build(orows, olv)
    return ix, o, st


def test_build_vector_l2_exact_adjacency():
    ix, o, st = build_both("vector_l2_ops", "f32", 6000, 64)
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    assert st["n_searched"] >= 5999 and st["n_rounds"] < st["n_searched"]      # really batched
    oc = o.build_counters()
    # phase A also runs for speculative searches that are later discarded, so the heuristic consults at least
    # the oracle's pairs.  A committed speculative search expands exactly the nodes the sequential search would
    # (that is what the commit rule proves); it may have evaluated a few elements the sequential run would have
    # found visited already, or vice versa -- all of them rejected either way -- so n_dist can differ marginally.
    assert st["n_pair"] >= oc["n_pair"]
    assert st["n_expand0"] == oc["n_expand0"] and st["n_expand_upper"] == oc["n_expand_upper"]
    assert abs(st["n_dist"] - oc["n_dist"]) <= 0.01 * oc["n_dist"]
    assert st["n_host_dist"] == 0 and st["n_pair_gpu"] > 0          # every distance of this build was computed on the GPU
    print("build stats", st)
    # the built graph is resident: scanning it matches the oracle's scan of its own graph
    q = make_rows("f32", 200, 64, 77)
    assert_same_results(ix.search(q, k=10, ef_search=40), o.search_batch(q, ef_search=40, k=10), "scan after build")
    ix.close()


@pytest.mark.parametrize("opclass,kind,dim", [("vector_ip_ops", "f32", 48), ("vector_cosine_ops", "f32", 24),
                                              ("vector_l1_ops", "f32", 17), ("halfvec_l2_ops", "f16", 40),
                                              ("halfvec_cosine_ops", "f16", 64), ("bit_hamming_ops", "bit", 256),
                                              ("bit_jaccard_ops", "bit", 96)])
def test_build_all_opclasses(opclass, kind, dim):
    ix, o, _ = build_both(opclass, kind, 1500, dim, m=8, efc=32)
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    ix.close()


@pytest.mark.parametrize("m,efc", [(2, 4), (4, 8), (24, 100), (40, 200)])
def test_build_option_range(m, efc):
    ix, o, _ = build_both("vector_l2_ops", "f32", 1200, 16, m=m, efc=efc)
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    ix.close()


def test_build_duplicates_and_zero_vectors():
    rng = np.random.default_rng(3)
    base = rng.standard_normal((300, 8)).astype(np.float32)
    rows = np.concatenate([base, base[:120], np.tile(base[5], (30, 1)), base[100:160]])
    rows = rows[rng.permutation(len(rows))]
    ix, o, st = build_both("vector_l2_ops", "f32", 0, 8, m=4, efc=16, rows=rows)
    assert st["n_duplicates"] > 100
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    ix.close()
    # cosine: zero vectors are not indexed (build.rs:433-435), parallel vectors are not byte-duplicates after normalising
    rows = np.concatenate([rows, np.zeros((3, 8), np.float32), 2 * rows[:10]])
    ix, o, _ = build_both("vector_cosine_ops", "f32", 0, 8, m=4, efc=16, rows=rows)
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    ix.close()


def test_build_levels_and_entry_point_changes():
    n = 800
    lv = np.zeros(n, np.int32)
    lv[[0, 5, 6, 40, 41, 300, 301, 302, 700]] = [1, 3, 3, 2, 7, 7, 9, 1, 12]
    ix, o, _ = build_both("vector_l2_ops", "f32", n, 12, m=6, efc=24, levels=lv)
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    ix.close()


def test_build_appends_across_calls_and_single_thread():
    rows = make_rows("f32", 2000, 20, 9)
    lv = D.hnsw_levels(2000, 8, 10)
    ix = P.HnswIndex("vector_l2_ops", 20, m=8, ef_construction=32, capacity=2000)
    ix.build(rows[:700], lv[:700], heap_tids=np.arange(700), host_threads=1)
    ix.build(rows[700:], lv[700:], heap_tids=np.arange(700, 2000), host_threads=3)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 20, m=8, ef_construction=32)
    o.build(rows, lv)
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    with pytest.raises(P.PgvrxError):
        ix.build(rows[:1], lv[:1])                         # beyond the handle's capacity
    ix.close()


def test_build_wide_rows_take_the_host_heuristic_path():
    # 2000-d fp32 rows: a CTA can stage only 25 rows, so layer-0 groups (33+ members) and the candidate lists fall back
    # to host distances for check_element_closer while the small upper-layer groups still use the GPU masks
    ix, o, st = build_both("vector_l2_ops", "f32", 400, 2000, m=16, efc=64)
    assert st["n_host_dist"] > 0
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    ix.close()


def test_build_tie_heavy_bits_and_many_duplicates():
    # 52-bit vectors: integer distances tie constantly and whole rows repeat far more than 10 times, so W's boundary,
    # the heuristic's `<=` and the replay rule's strict `<` are all exercised on equal distances
    rows = make_rows("bit", 3000, 52, 19)
    rows[::7] = rows[3]
    ix, o, st = build_both("bit_hamming_ops", "bit", 0, 52, m=8, efc=32, rows=rows)
    assert st["n_duplicates"] > 10
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    ix.close()


def test_build_20k_rows_many_commits_per_round():
    ix, o, st = build_both("vector_l2_ops", "f32", 20000, 24, m=12, efc=48)
    assert st["n_rounds"] * 3 < 20000                       # several commits per round on average
    assert_same_graph(ix.export_graph(), o.export_graph(), o)
    ix.close()
