"""The oracle's restatement of amgettuple's iterative scans, aminsert and the duplicate rule against the reference's
remaining TAP tests of the path: tests/t/043 and 044 (iterative scans under a filter), 016 (an index that only ever
saw aminsert), 023 and 027 (duplicates for bit / halfvec, built and inserted).  Same row counts, distributions,
GUC values, LIMITs and thresholds as the Perl; the executor's part (apply the WHERE clause to what the index scan
emits, stop at LIMIT) is the few lines of `filtered` below.  The planner tests (017, 039) and WAL (010) are outside
the path."""
import numpy as np
import pytest

import oracle_ffi as O
from helpers import normalize_rows
from pgvector_rx_b200 import datasets as D


def filtered(o, q, keep, limit, mode, ef_search=40, max_scan_tuples=20000):
    """SELECT ... WHERE keep(i) ORDER BY v <op> q LIMIT limit over the index scan: heap TIDs in emission order, filtered,
    cut at LIMIT (xs_recheckorderby is false, scan.rs:808-811, so the executor does not re-sort)."""
    tids, _, _, _ = o.scan_stream(q, ef_search=ef_search, iterative=mode, max_scan_tuples=max_scan_tuples)
    hits = [int(t) for t in tids if keep(int(t))]
    return hits[:limit]


@pytest.mark.parametrize("opclass", ["vector_l2_ops", "vector_cosine_ops"])
def test_tap_044_iterative_scan_recall(opclass):
    # tests/t/044_hnsw_iterative_scan_recall.pl: 50000 rows of random() in 3-d, 20 queries, WHERE i % c = 0 for
    # c in (50, 500), LIMIT 20, ef_search = 40, strict_order and relaxed_order: recall >= 0.99 against the rows
    # within the 20th filtered distance (:77-83, :96-97)
    n, limit = 50000, 20
    rng = np.random.default_rng(44)
    rows = rng.random((n, 3)).astype(np.float32)
    queries = rng.random((20, 3)).astype(np.float32)
    if opclass == "vector_cosine_ops":
        stored, metric = normalize_rows(O.ELEM_F32, 3, rows), O.NEG_IP            # build.rs:426-438
        assert len(stored) == n
    else:
        stored, metric = rows, O.L2SQ
    o = O.Oracle(O.ELEM_F32, metric, 3)
    o.build(stored, D.hnsw_levels(n, 16, 45))
    i_of = np.arange(1, n + 1)                                                    # generate_series(1, 50000), heap TID t = row t
    r64 = rows.astype(np.float64)
    for c in (50, 500):
        for mode in (O.ITER_STRICT, O.ITER_RELAXED):
            correct = total = 0
            for q in queries:
                q64 = q.astype(np.float64)
                if metric == O.L2SQ:
                    d = ((r64 - q64) ** 2).sum(-1)
                    qq = q
                else:
                    d = 1.0 - (r64 @ q64) / (np.linalg.norm(r64, axis=1) * np.linalg.norm(q64))
                    qq = O.normalize(O.ELEM_F32, 3, q)[1]                         # scan.rs:749-754
                dmax = np.sort(d[i_of % c == 0])[limit - 1]
                expected = set(i_of[d <= dmax * (1 + 1e-12)].tolist())
                got = filtered(o, qq, lambda t: (t + 1) % c == 0, limit, mode)
                correct += sum((t + 1) in expected for t in got)
                total += limit
            assert correct / total >= 0.99, (opclass, c, mode, correct / total)


def test_tap_043_iterative_scan_max_scan_tuples():
    # tests/t/043_hnsw_iterative_scan.pl: 100000 rows, relaxed_order, WHERE i % 10000 = 0 LIMIT 11.  With
    # max_scan_tuples = 100000 all 10 matching rows come back (:28-36); with 30000 / 50000 / 70000 the average
    # count over 20 queries is within 2 of max_scan_tuples / 10000 (:38-58)
    n = 100000
    rng = np.random.default_rng(43)
    rows = rng.random((n, 3)).astype(np.float32)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    o.build(rows, D.hnsw_levels(n, 16, 46))
    keep = lambda t: (t + 1) % 10000 == 0
    assert len(filtered(o, rows[0], keep, 11, O.ITER_RELAXED, max_scan_tuples=100000)) == 10
    for max_tuples in (30000, 50000, 70000):
        counts = [len(filtered(o, rows[i], keep, 11, O.ITER_RELAXED, max_scan_tuples=max_tuples)) for i in range(20)]
        avg = sum(counts) / 20
        assert max_tuples / 10000 - 2 < avg < max_tuples / 10000 + 2, (max_tuples, counts)


def test_tap_016_inserts_only():
    # tests/t/016_hnsw_inserts.pl without the concurrency: an index created empty and filled by aminsert alone.
    # 10 rows of 1900 dimensions: an ordered scan from one of them returns all 10 (:34-38); 1000 rows with
    # ef_search = 1000: at least 997 come back (:55-60)
    dim = 1900
    rng = np.random.default_rng(16)
    for trial in range(3):
        rows = rng.random((10, dim)).astype(np.float32)
        o = O.Oracle(O.ELEM_F32, O.L2SQ, dim)
        o.insert(rows, D.hnsw_levels(10, 16, 160 + trial))
        tids, _, _, _ = o.scan_stream(rows[0], ef_search=40)
        assert sorted(tids.tolist()) == list(range(10))
    rows = rng.random((1000, dim)).astype(np.float32)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, dim)
    o.insert(rows, D.hnsw_levels(1000, 16, 170))
    tids, _, _, _ = o.scan_stream(rows[0], ef_search=1000)
    assert len(tids) >= 997 and len(set(tids.tolist())) == len(tids)


@pytest.mark.parametrize("elem,metric,row", [(O.ELEM_BIT, O.HAMMING, np.array([[0b11100000]], np.uint8)),
                                             (O.ELEM_F16, O.L2SQ, None)])
@pytest.mark.parametrize("path", ["build", "insert"])
def test_tap_023_027_duplicates(elem, metric, row, path):
    # tests/t/023_hnsw_bit_duplicates.pl, 027_hnsw_halfvec_duplicates.pl: 20 copies of one value, ef_search = 1:
    # the scan returns exactly 10 rows (one element's heap TIDs), after CREATE INDEX (:28-29) and after
    # TRUNCATE + 20 single-row INSERTs (:30-32)
    if row is None:
        row = np.array([[O.lib().orc_f32_to_half(1.0)] * 3], np.uint16)
    rows = np.repeat(row, 20, axis=0)
    o = O.Oracle(elem, metric, 3)
    lv = D.hnsw_levels(20, 16, 23)
    if path == "build":
        o.build(rows, lv)
    else:
        for i in range(20):
            o.insert(rows[i:i + 1], lv[i:i + 1], base_tid=i)
    assert o.n == 2                                                   # 10 + 10 heap TIDs (hnsw_constants.rs:85)
    tids, _, _, _ = o.scan_stream(row[0], ef_search=1)
    assert len(tids) == 10 and len(set(tids.tolist())) == 10
