"""The two tie policies of the oracle (SURVEY 8a note T): RUST_HEAP restates std BinaryHeap's
internal order (what the reference does, unpinned by any reference test); TOTAL_ORDER breaks
equal distances by element id (what the GPU does).  They must agree wherever distances are
distinct, and be tie-equivalent -- same distances, possibly different ids among equals -- otherwise,
which is exactly how the reference's own bit tests compare (tests/t/020_hnsw_bit_build_recall.pl:85-91)."""
import numpy as np

import oracle_ffi as O
from helpers import make_rows, oracle_index


def test_tie_free_data_modes_are_identical():
    rows = make_rows("f32", 4000, 32, 1)
    q = make_rows("f32", 300, 32, 2)
    a = oracle_index("vector_l2_ops", 32, rows, tie=O.TIE_RUST_HEAP)
    b = oracle_index("vector_l2_ops", 32, rows, tie=O.TIE_TOTAL_ORDER)
    ga, gb = a.export_graph(), b.export_graph()
    for k in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(ga[k], gb[k]), k            # identical adjacency after build
    assert ga["entry"] == gb["entry"]
    for ef in (1, 10, 40, 100):
        ra, rb = a.search_batch(q, ef, 10), b.search_batch(q, ef, 10)
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2])
    ta = a.scan_stream(q[0], 5, O.ITER_RELAXED, max_scan_tuples=200)
    tb = b.scan_stream(q[0], 5, O.ITER_RELAXED, max_scan_tuples=200)
    assert np.array_equal(ta[0], tb[0])


def test_tie_heavy_data_modes_are_tie_equivalent():
    rows = make_rows("bit", 3000, 64, 1)
    q = make_rows("bit", 300, 64, 2)
    b = oracle_index("bit_hamming_ops", 64, rows, tie=O.TIE_TOTAL_ORDER)
    a = O.Oracle(O.ELEM_BIT, O.HAMMING, 64, tie_mode=O.TIE_RUST_HEAP)
    a.import_graph(b.export_graph())                       # same graph, only the scan's tie policy differs
    ra, rb = a.search_batch(q, 40, 10), b.search_batch(q, 40, 10)
    same_dists = (ra[1] == rb[1]).all(axis=1).mean()
    assert same_dists >= 0.9                               # equal distances, ids may differ among ties
    assert abs(ra[1][:, 9].mean() - rb[1][:, 9].mean()) <= 0.01 * rb[1][:, 9].mean()
    # every returned id really has the distance reported for it, in both modes
    for r in (ra, rb):
        for i in range(20):
            for j in range(10):
                assert O.distance(O.ELEM_BIT, O.HAMMING, 64, q[i], rows[r[0][i, j]]) == r[1][i, j]


def test_live_index_paths_agree_across_tie_modes_on_tie_free_data():
    # aminsert and ambulkdelete restated under both tie policies: with distinct distances the reference's own heap
    # order (RUST_HEAP) and the GPU's (distance, id) order must leave the same graph
    from pgvector_rx_b200 import datasets as D
    rows = make_rows("f32", 3000, 16, 3)
    lv = D.hnsw_levels(3000, 8, 4)
    out = []
    for tie in (O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER):
        o = O.Oracle(O.ELEM_F32, O.L2SQ, 16, m=8, ef_construction=32, tie_mode=tie)
        o.build(rows[:1500], lv[:1500])
        o.insert(rows[1500:], lv[1500:], base_tid=1500)
        g1 = o.export_graph()
        o.bulk_delete(np.arange(0, 3000, 3))
        out.append((g1, o.export_graph(), o.element_flags(), o.entry))
    (a1, a2, fa, ea), (b1, b2, fb, eb) = out
    for k in ("level", "adj0", "upper_base", "adj_up"):
        assert np.array_equal(a1[k], b1[k]), ("after insert", k)
        assert np.array_equal(a2[k], b2[k]), ("after vacuum", k)
    assert np.array_equal(fa, fb) and ea == eb and a1["entry"] == b1["entry"]
