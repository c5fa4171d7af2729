"""GPU parity of the iterative index scan (scan.rs:538-577, 794-875): the stream of elements a
query emits across get_scan_items / resume_scan_items rounds -- relaxed and strict order, the
hnsw.max_scan_tuples drain, consumers that stop early -- against the oracle's amgettuple stream, element
for element with bit-identical distances and the same work counters."""
import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import gpu_index, make_rows, oracle_index

pytestmark = pytest.mark.gpu


def oracle_streams(o, qn, ef, mode, max_tuples, limit):
    out, ctr = [], dict(n_dist=0, n_expand0=0, n_expand_upper=0, n_visited=0)
    for q in qn:
        tids, eids, ds, c = o.scan_stream(q, ef_search=ef, iterative=mode, max_scan_tuples=max_tuples,
                                          max_emit=(1 << 30) if limit is None else limit)
        out.append((eids, ds.astype(np.float32)))
        for k in ctr:
            ctr[k] += c[k]
    return out, ctr


def run_stream_case(opclass, kind, n, dim, nq, ef, mode, max_tuples=20000, limit=None, m=16, efc=64, seed=71,
                    check_counters=True):
    rows = make_rows(kind, n, dim, seed)
    o = oracle_index(opclass, dim, rows, m=m, efc=efc)
    ix = gpu_index(opclass, dim, o.export_graph(), m=m, efc=efc)
    q = make_rows(kind, nq, dim, seed + 1)
    qn = ix.prepare_queries(q)
    sc = ix.scan_begin(q, ef_search=ef, iterative_scan=mode, max_scan_tuples=max_tuples)
    got_i, got_d = sc.stream(limit)
    want, ctr = oracle_streams(o, qn, ef, mode, max_tuples, limit)
    for i in range(nq):
        wi, wd = want[i]
        assert len(got_i[i]) == len(wi), "query %d: %d elements, oracle %d" % (i, len(got_i[i]), len(wi))
        bad = np.nonzero(got_i[i] != wi)[0]
        assert len(bad) == 0, "query %d differs first at position %d: %s vs %s" % (
            i, bad[0], got_i[i][bad[0]:bad[0] + 5], wi[bad[0]:bad[0] + 5])
        assert np.array_equal(got_d[i], wd)
    if check_counters and limit is None:
        # same traversal => same work; only comparable when both sides ran the stream to exhaustion
        for k in ("n_dist", "n_expand0", "n_expand_upper", "n_visited"):
            assert sc.totals[k] == ctr[k], (k, sc.totals[k], ctr[k])
    sc.close()
    ix.close()
    return got_i


@pytest.mark.parametrize("mode", [P.ITER_RELAXED, P.ITER_STRICT])
def test_full_stream_vector_l2(mode):
    got = run_stream_case("vector_l2_ops", "f32", 3000, 32, 24, 40, mode)
    # an exhausted relaxed scan reaches (nearly) every element of a connected graph
    assert min(len(g) for g in got) > 2500


@pytest.mark.parametrize("ef", [1, 3, 100])
def test_full_stream_ef_range(ef):
    # hnsw_vector.out:107-124: with ef_search = 1 an iterative scan still returns every row
    run_stream_case("vector_l2_ops", "f32", 800, 16, 12, ef, P.ITER_RELAXED, m=8, efc=32)
    run_stream_case("vector_l2_ops", "f32", 800, 16, 12, ef, P.ITER_STRICT, m=8, efc=32)


@pytest.mark.parametrize("mode", [P.ITER_RELAXED, P.ITER_STRICT])
def test_max_scan_tuples_drains_discarded(mode):
    # after max_scan_tuples elements the remaining discarded candidates are returned nearest-first without searching
    run_stream_case("vector_l2_ops", "f32", 3000, 32, 16, 40, mode, max_tuples=120)
    run_stream_case("vector_l2_ops", "f32", 3000, 32, 16, 40, mode, max_tuples=1)


@pytest.mark.parametrize("limit", [1, 25, 41, 333])
def test_consumer_stops_early(limit):
    run_stream_case("vector_l2_ops", "f32", 4000, 48, 40, 40, P.ITER_RELAXED, limit=limit)


def test_iterative_off_equals_plain_scan():
    rows = make_rows("f32", 2000, 24, 5)
    o = oracle_index("vector_l2_ops", 24, rows)
    ix = gpu_index("vector_l2_ops", 24, o.export_graph())
    q = make_rows("f32", 50, 24, 6)
    sc = ix.scan_begin(q, ef_search=40, iterative_scan=P.ITER_OFF)
    ids, dists, counts, done = sc.next()
    assert done.all()
    want = ix.search(q, k=40, ef_search=40)
    assert np.array_equal(ids, want[0]) and np.array_equal(dists, want[1]) and np.array_equal(counts, want[2])
    ids2, _, counts2, done2 = sc.next()
    assert (counts2 == 0).all() and done2.all() and (ids2 == -1).all()
    sc.close()
    ix.close()


def test_c4_shape_ip_relaxed_with_filter():
    # BASELINE config 4 shape (scaled down): vector_ip_ops, 96-d, relaxed_order, a consumer that keeps
    # id % c == 0 and stops after k' hits (tests/t/044: filtered scans are what iterative scans are for)
    n, dim, nq, kp = 6000, 96, 30, 10
    rows = make_rows("f32", n, dim, 7)
    o = oracle_index("vector_ip_ops", dim, rows)
    ix = gpu_index("vector_ip_ops", dim, o.export_graph())
    q = make_rows("f32", nq, dim, 8)
    for c in (50, 500):
        sc = ix.scan_begin(q, ef_search=40, iterative_scan=P.ITER_RELAXED, max_scan_tuples=20000)
        hits = [[] for _ in range(nq)]
        done = np.zeros(nq, bool)
        while True:
            want = ~done & np.array([len(h) < kp for h in hits])
            if not want.any():
                break
            ids, dists, counts, d = sc.next(want.astype(np.uint8))
            for i in np.nonzero(want)[0]:
                for e in ids[i, :counts[i]]:
                    if e % c == 0 and len(hits[i]) < kp:
                        hits[i].append(int(e))
            done |= d
        sc.close()
        for i in range(nq):
            _, eids, _, _ = o.scan_stream(q[i], ef_search=40, iterative=O.ITER_RELAXED, max_scan_tuples=20000)
            want_hits = [int(e) for e in eids if e % c == 0][:kp]
            assert hits[i] == want_hits, (c, i, hits[i], want_hits)
    ix.close()


@pytest.mark.parametrize("opclass,kind,dim", [("bit_hamming_ops", "bit", 256), ("halfvec_cosine_ops", "f16", 64),
                                              ("vector_l1_ops", "f32", 20)])
def test_other_opclasses(opclass, kind, dim):
    # bit distances tie constantly: evicted-but-unexpanded elements appear twice in `discarded`, like the reference
    run_stream_case(opclass, kind, 1500, dim, 10, 20, P.ITER_RELAXED, m=8, efc=32)
    run_stream_case(opclass, kind, 1500, dim, 10, 20, P.ITER_STRICT, m=8, efc=32, max_tuples=200)


def test_tiny_and_empty_indexes_and_errors():
    for n in (1, 2, 9):
        run_stream_case("vector_l2_ops", "f32", n, 4, 5, 40, P.ITER_RELAXED, m=4, efc=16)
    ix = P.HnswIndex("vector_l2_ops", 4, capacity=4)
    with pytest.raises(P.PgvrxError) as e:
        ix.scan_begin(np.zeros((1, 4), np.float32))
    assert e.value.code == P.ERR_STATE
    ix.load_graph(np.zeros((0, 4), np.float32), np.zeros(0, np.uint8), np.zeros((0, 32), np.int32),
                  np.zeros(0, np.int32), np.zeros((0, 16), np.int32), -1)
    sc = ix.scan_begin(np.ones((2, 4), np.float32))
    ids, dists, counts, done = sc.next()
    assert (counts == 0).all() and done.all()
    sc.close()
    for kw in (dict(ef_search=0), dict(ef_search=1001), dict(iterative_scan=7), dict(max_scan_tuples=0)):
        with pytest.raises(P.PgvrxError) as e:
            ix.scan_begin(np.ones((1, 4), np.float32), **kw)
        assert e.value.code == P.ERR_INVALID
    ix.close()


def test_heap_tid_fanout_with_duplicates():
    # duplicated rows share one element with up to 10 heap TIDs, emitted from the back of its list (scan.rs:798)
    rng = np.random.default_rng(5)
    base = rng.standard_normal((200, 8)).astype(np.float32)
    rows = np.concatenate([base, base[:60], base[:60], np.tile(base[7], (15, 1))])
    rows = rows[rng.permutation(len(rows))]
    lv = np.zeros(len(rows), np.int32)
    lv[::17] = 1
    ix = P.HnswIndex("vector_l2_ops", 8, m=4, ef_construction=16, capacity=len(rows))
    ix.build(rows, lv)                                   # heap TID of a row = its row number
    g = ix.export_graph()
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 8, m=4, ef_construction=16, tie_mode=O.TIE_TOTAL_ORDER)
    o.build(rows, lv)
    q = base[[7, 3, 150]] + 0.01
    for mode in (P.ITER_OFF, P.ITER_RELAXED, P.ITER_STRICT):
        sc = ix.scan_begin(q, ef_search=10, iterative_scan=mode)
        ids, _ = sc.stream(None)
        sc.close()
        for i in range(len(q)):
            tids, eids, _, _ = o.scan_stream(q[i], ef_search=10, iterative=mode, max_scan_tuples=20000)
            # the oracle reports the element of every emitted TID: each element repeats once per heap TID
            assert np.array_equal(np.repeat(ids[i], g["heap_tid_counts"][ids[i]]), eids)
            got = P.heap_tid_stream(ids[i], g["heap_tids"], g["heap_tid_counts"])
            assert np.array_equal(got.astype(np.int64), tids)
    assert g["heap_tid_counts"].max() == 10
    ix.close()


@pytest.mark.parametrize("mode,rounds,max_tuples", [(P.ITER_RELAXED, 4, 20000), (P.ITER_STRICT, 7, 20000), (P.ITER_RELAXED, 16, 600)])
def test_several_refills_per_call_give_the_same_stream(mode, rounds, max_tuples):
    # pgvrx_scan_next_rounds: `rounds` refills in one launch == `rounds` calls of pgvrx_scan_next (also across the switch to
    # the max_scan_tuples drain, and when a query runs dry in the middle of a call)
    n, dim, nq, ef = 2500, 24, 20, 12
    rows = make_rows("f32", n, dim, 91)
    o = oracle_index("vector_l2_ops", dim, rows, m=8, efc=32)
    ix = gpu_index("vector_l2_ops", dim, o.export_graph(), m=8, efc=32)
    q = make_rows("f32", nq, dim, 92)
    one = ix.scan_begin(q, ef_search=ef, iterative_scan=mode, max_scan_tuples=max_tuples)
    want_i, want_d = one.stream(None)
    many = ix.scan_begin(q, ef_search=ef, iterative_scan=mode, max_scan_tuples=max_tuples)
    got_i = [[] for _ in range(nq)]
    got_d = [[] for _ in range(nq)]
    done = np.zeros(nq, bool)
    calls = 0
    while not done.all():
        ids, dists, counts, d = many.next((~done).astype(np.uint8), rounds=rounds)
        assert ids.shape == (nq, rounds, ef) and counts.shape == (nq, rounds)
        for i in np.nonzero(~done)[0]:
            for r in range(rounds):
                got_i[i].extend(ids[i, r, :counts[i, r]].tolist())
                got_d[i].extend(dists[i, r, :counts[i, r]].tolist())
                assert (ids[i, r, counts[i, r]:] == -1).all()
        done |= d
        calls += 1
    for i in range(nq):
        assert got_i[i] == list(want_i[i]), i
        assert np.array_equal(np.asarray(got_d[i], np.float32), want_d[i]), i
    for k in ("n_dist", "n_expand0", "n_visited"):
        assert many.totals[k] == one.totals[k], k
    assert calls * rounds >= one.totals["rounds"] > (calls - 1) * rounds
    one.close()
    many.close()
    ix.close()
