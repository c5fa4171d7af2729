"""GPU parity of VACUUM (ambulkdelete, src/index/vacuum.rs:816-842): pgvrx_index_vacuum against the
oracle's restatement of the three passes -- the same dead heap TIDs must leave the same graph (every repaired list, the
entry point), the same element flags and heap TIDs, although the GPU runs the skip-set repair searches in speculative
batches.  Scans afterwards must agree too."""
import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import assert_same_results, make_rows
from pgvector_rx_b200 import datasets as D

pytestmark = pytest.mark.gpu


def same_state(ix, o, what):
    g, w = ix.export_graph(), o.export_graph()
    assert g["n"] == w["n"] and g["entry"] == w["entry"], (what, g["entry"], w["entry"])
    bad = np.nonzero((g["adj0"] != w["adj0"]).any(axis=1))[0]
    assert len(bad) == 0, "%s: %d layer-0 lists differ, first %d:\n gpu %s\n orc %s" % (
        what, len(bad), bad[0], g["adj0"][bad[0]], w["adj0"][bad[0]])
    assert np.array_equal(g["adj_up"], w["adj_up"]), what
    assert np.array_equal(ix.element_flags(), o.element_flags()), what
    assert np.array_equal(g["rows"].view(np.uint8), w["rows"].view(np.uint8)), what
    for i in range(0, o.n, max(1, o.n // 200)):
        t = o.heap_tids(i)
        assert g["heap_tid_counts"][i] == len(t) and g["heap_tids"][i, :len(t)].astype(np.int64).tolist() == t.tolist()


def run_case(opclass, kind, n, dim, m, efc, rounds, seed=51, kill_entry=False):
    elem, metric, norm = P.OPCLASSES[opclass]
    rows = make_rows(kind, n, dim, seed)
    lv = D.hnsw_levels(n, m, seed + 1)
    o = O.Oracle(elem, metric, dim, m=m, ef_construction=efc, tie_mode=O.TIE_TOTAL_ORDER)
    o.build(rows, lv)
    ix = P.HnswIndex(opclass, dim, m=m, ef_construction=efc, capacity=n)
    ix.build(rows, lv, heap_tids=np.arange(n, dtype=np.uint64))
    rng = np.random.default_rng(seed)
    alive = np.arange(n)
    q = make_rows(kind, 100, dim, seed + 2)
    for r, frac in enumerate(rounds):
        dead = rng.permutation(alive)[: int(len(alive) * frac)]
        if kill_entry and r == 0 and o.entry not in dead:
            dead = np.append(dead, o.entry)
        alive = np.setdiff1d(alive, dead)
        want_removed = o.bulk_delete(dead)
        st = ix.vacuum(dead)
        assert st["n_heap_tids_removed"] == want_removed == len(dead)
        assert st["n_elements_deleted"] == len(dead) and st["n_rounds"] < max(8, st["n_elements_repaired"])
        same_state(ix, o, "%s vacuum %d" % (opclass, r))
        if len(alive):
            assert_same_results(ix.search(q, k=10, ef_search=40), o.search_batch(q, ef_search=40, k=10), "scan after vacuum %d" % r)
    ix.close()


def test_vacuum_l2_two_passes():
    run_case("vector_l2_ops", "f32", 3000, 16, 6, 24, rounds=(0.3, 0.3))


def test_vacuum_deleting_the_entry_point():
    run_case("vector_l2_ops", "f32", 1500, 12, 4, 16, rounds=(0.2,), kill_entry=True)


def test_vacuum_bit_hamming_ties():
    run_case("bit_hamming_ops", "bit", 2000, 64, 6, 24, rounds=(0.4,))


def test_vacuum_everything_then_nothing_is_found():
    run_case("vector_l2_ops", "f32", 400, 8, 4, 16, rounds=(1.0,))
