"""The oracle's restatement of CREATE INDEX + scan against the reference's TAP build-recall tests (tests/t/012, 020, 024):
10000 seeded rows of the same distributions, default m = 16 / ef_construction = 64, 20 queries, LIMIT 20, the ef_search
each test sets, the same recall definition and the same thresholds.  Both tie policies are run: RUST_HEAP is what the
reference does, TOTAL_ORDER what the GPU does."""
import numpy as np
import pytest

import oracle_ffi as O
from pgvector_rx_b200 import datasets as D
from test_oracle_live_index import brute, recall, tie_aware_recall, half_rows


@pytest.mark.parametrize("tie", [O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER])
@pytest.mark.parametrize("metric,kind,min_recall", [(O.L2SQ, "l2", 0.99), (O.NEG_IP, "ip", 0.97), (O.L1, "l1", 0.99)])
def test_tap_012_vector_build_recall(metric, kind, min_recall, tie):
    # tests/t/012_hnsw_vector_build_recall.pl: 3-d rows of random()*random(); default ef_search = 40
    rng = np.random.default_rng(12)
    rows = (rng.random((10000, 3)) * rng.random((10000, 3))).astype(np.float32)
    q = rng.random((20, 3)).astype(np.float32)
    o = O.Oracle(O.ELEM_F32, metric, 3, tie_mode=tie)
    o.build(rows, D.hnsw_levels(len(rows), 16, 13))
    ids, _, _ = o.search_batch(q, ef_search=40, k=20)
    r64, q64 = rows.astype(np.float64), q.astype(np.float64)
    if kind == "l1":
        d = np.abs(r64[None] - q64[:, None]).sum(-1)
        expected = np.argsort(d, axis=1, kind="stable")[:, :20]
    else:
        expected = brute(r64, q64, 20, kind)
    first_row = np.array([int(o.heap_tids(e)[0]) for e in range(o.n)])
    assert recall(expected, [first_row[r[r >= 0]].tolist() for r in ids]) >= min_recall


@pytest.mark.parametrize("metric,min_recall", [(O.HAMMING, 0.98), (O.JACCARD, 0.95)])
def test_tap_020_bit_build_recall(metric, min_recall):
    # tests/t/020_hnsw_bit_build_recall.pl: random 52-bit rows, hnsw.ef_search = 100 (:22,:31), tie-aware expected set
    tab = np.array([bin(i).count("1") for i in range(256)], np.float64)
    rows = D.random_bits(10000, 52, 320)
    q = D.random_bits(20, 52, 321)
    rows[:, -1] &= 0xF0
    q[:, -1] &= 0xF0
    o = O.Oracle(O.ELEM_BIT, metric, 52, tie_mode=O.TIE_RUST_HEAP)
    o.build(rows, D.hnsw_levels(len(rows), 16, 322))
    ids, _, _ = o.search_batch(q, ef_search=100, k=20)
    if metric == O.HAMMING:
        d = tab[rows[None, :, :] ^ q[:, None, :]].sum(-1)
    else:
        ab = tab[rows[None, :, :] & q[:, None, :]].sum(-1)
        aa, bb = tab[q].sum(-1)[:, None], tab[rows].sum(-1)[None, :]
        d = np.where(ab == 0, 1.0, 1.0 - ab / np.maximum(aa + bb - ab, 1))
    first_row = np.array([int(o.heap_tids(e)[0]) for e in range(o.n)])
    ids_rows = np.where(ids >= 0, first_row[np.maximum(ids, 0)], -1)
    assert tie_aware_recall(d, ids_rows, 20) >= min_recall


@pytest.mark.parametrize("metric,kind", [(O.L2SQ, "l2"), (O.NEG_IP, "ip"), (O.L1, "l1")])
def test_tap_024_halfvec_build_recall(metric, kind):
    # tests/t/024_hnsw_halfvec_build_recall.pl: 10-d rows of 2*random()*random(), >= 0.98 at the default ef_search
    rng = np.random.default_rng(24)
    rows = half_rows(2 * rng.random((10000, 10)) * rng.random((10000, 10)))
    q = half_rows(rng.random((20, 10)))
    o = O.Oracle(O.ELEM_F16, metric, 10, tie_mode=O.TIE_RUST_HEAP)
    o.build(rows, D.hnsw_levels(len(rows), 16, 25))
    ids, _, _ = o.search_batch(q, ef_search=40, k=20)
    r64, q64 = rows.view(np.float16).astype(np.float64), q.view(np.float16).astype(np.float64)
    if kind == "l2":
        d = ((r64[None] - q64[:, None]) ** 2).sum(-1)
    elif kind == "ip":
        d = -(r64[None] * q64[:, None]).sum(-1)
    else:
        d = np.abs(r64[None] - q64[:, None]).sum(-1)
    first_row = np.array([int(o.heap_tids(e)[0]) for e in range(o.n)])
    ids_rows = np.where(ids >= 0, first_row[np.maximum(ids, 0)], -1)
    assert tie_aware_recall(d, ids_rows, 20) >= 0.98
