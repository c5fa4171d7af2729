"""GPU parity of the exact paths: flat kNN, exact re-rank and the k-way top-k merge, through
the C ABI, against the oracle's brute force (sequential scan sorted by the operator value, ties by id).
Ids identical, distances bit-identical."""
import os

import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import make_rows, normalize_rows

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def prepared(opclass, kind, n, dim, nq, seed):
    """rows / queries as the index and the oracle see them (cosine opclasses: unit-normalised on both sides)."""
    elem, metric, norm = P.OPCLASSES[opclass]
    rows = make_rows(kind, n, dim, seed)
    q = make_rows(kind, nq, dim, seed + 1)
    if norm:
        rows, q = normalize_rows(elem, dim, rows), normalize_rows(elem, dim, q)
    return elem, metric, rows, q


def flat_case(opclass, kind, n, dim, nq, k, seed=31):
    elem, metric, rows, q = prepared(opclass, kind, n, dim, nq, seed)
    # rows are already normalised: load them through a non-normalising opclass of the same metric
    raw = opclass.replace("cosine", "ip")
    ix = P.HnswIndex(raw, dim, capacity=max(len(rows), 1))
    ix.load_rows(rows)
    ids, dists, st = ix.flat_topk(q, k=k, with_stats=True)
    want_ids, want_d = O.flat_topk(elem, metric, dim, rows, q, k, nthreads=THREADS)
    bad = np.nonzero((ids != want_ids).any(axis=1))[0]
    assert len(bad) == 0, "%s: ids differ for %d queries, first %d: %s vs %s (%s vs %s)" % (
        opclass, len(bad), bad[0], ids[bad[0]], want_ids[bad[0]], dists[bad[0]], want_d[bad[0]])
    assert np.array_equal(dists, want_d)
    assert st["n_pairs"] == len(q) * len(rows) and st["n_launches"] >= 1
    ix.close()
    return st


def test_flat_c2_shape_l2():
    flat_case("vector_l2_ops", "f32", 20000, 128, 500, 10)


@pytest.mark.parametrize("opclass", ["vector_ip_ops", "vector_cosine_ops", "vector_l1_ops"])
def test_flat_vector_metrics(opclass):
    flat_case(opclass, "f32", 5000, 96, 100, 10)


@pytest.mark.parametrize("opclass", ["halfvec_l2_ops", "halfvec_ip_ops", "halfvec_cosine_ops", "halfvec_l1_ops"])
def test_flat_halfvec_metrics(opclass):
    # C3 shape: 768-d halves span three 512-byte dimension chunks
    flat_case(opclass, "f16", 3000, 768, 70, 10)


@pytest.mark.parametrize("opclass,dim", [("bit_hamming_ops", 1024), ("bit_jaccard_ops", 520), ("bit_hamming_ops", 52)])
def test_flat_bit_metrics_tie_heavy(opclass, dim):
    # integer distances tie constantly; ties are ordered by row id on both sides
    flat_case(opclass, "bit", 6000, dim, 100, 20)


@pytest.mark.parametrize("dim", [1, 3, 5, 17, 130, 300, 2000])
def test_flat_ragged_dims(dim):
    flat_case("vector_l2_ops", "f32", 700, dim, 37, 7)


@pytest.mark.parametrize("n,k", [(1, 5), (7, 10), (127, 128), (129, 100), (1000, 1), (1000, 300)])
def test_flat_small_tables_and_k(n, k):
    flat_case("vector_l2_ops", "f32", n, 16, 33, k)


def test_flat_duplicates_order_by_id():
    rows = np.tile(np.arange(8, dtype=np.float32)[None, :], (300, 1))
    rows[::3] += 1.0
    ix = P.HnswIndex("vector_l2_ops", 8, capacity=300)
    ix.load_rows(rows)
    q = rows[:4]
    ids, dists = ix.flat_topk(q, k=50)
    wi, wd = O.flat_topk(O.ELEM_F32, O.L2SQ, 8, rows, q, 50)
    assert np.array_equal(ids, wi) and np.array_equal(dists, wd)
    ix.close()


def test_flat_empty_table_and_errors():
    ix = P.HnswIndex("vector_l2_ops", 4, capacity=8)
    with pytest.raises(P.PgvrxError) as e:
        ix.flat_topk(np.zeros((1, 4), np.float32))
    assert e.value.code == P.ERR_STATE
    ix.load_rows(np.zeros((0, 4), np.float32))
    ids, dists = ix.flat_topk(np.ones((3, 4), np.float32), k=5)
    assert (ids == -1).all() and np.isinf(dists).all()
    ix.load_rows(np.eye(4, dtype=np.float32))
    for k in (0, 2000):
        with pytest.raises(P.PgvrxError):
            ix.flat_topk(np.ones((1, 4), np.float32), k=k)
    ix.close()


def test_flat_works_on_a_loaded_graph_handle():
    from helpers import gpu_index, oracle_index
    rows = make_rows("f32", 3000, 64, 5)
    o = oracle_index("vector_l2_ops", 64, rows)
    ix = gpu_index("vector_l2_ops", 64, o.export_graph())
    q = make_rows("f32", 64, 64, 6)
    ids, dists = ix.flat_topk(q, k=10)
    wi, wd = O.flat_topk(O.ELEM_F32, O.L2SQ, 64, rows, q, 10, nthreads=THREADS)
    assert np.array_equal(ids, wi) and np.array_equal(dists, wd)
    # recall of the graph scan against the exact result, computed entirely on the GPU side
    got = ix.search(q, k=10, ef_search=40)[0]
    hits = sum(len(set(got[i]) & set(ids[i])) for i in range(len(q)))
    assert hits / (10.0 * len(q)) > 0.5
    ix.close()


@pytest.mark.parametrize("opclass,kind,dim", [("vector_l2_ops", "f32", 128), ("halfvec_ip_ops", "f16", 768),
                                              ("bit_hamming_ops", "bit", 1024), ("vector_l1_ops", "f32", 2000)])
def test_rerank_matches_brute_force_on_the_candidates(opclass, kind, dim):
    elem, metric, rows, q = prepared(opclass, kind, 4000, dim, 50, 41)
    ix = P.HnswIndex(opclass, dim, capacity=len(rows))
    ix.load_rows(rows)
    rng = np.random.default_rng(9)
    nc, k = 77, 10
    cand = np.stack([rng.choice(len(rows), nc, replace=False) for _ in range(len(q))]).astype(np.int32)
    cand[:, 5] = -1                       # holes are skipped
    cand[3, :] = -1                       # a query without candidates returns nothing
    ids, dists = ix.rerank(q, cand, k=k)
    for i in range(len(q)):
        c = np.sort(cand[i][cand[i] >= 0])     # ties are ordered by element id: rank the subset in id order
        if len(c) == 0:
            assert (ids[i] == -1).all() and np.isinf(dists[i]).all()
            continue
        wi, wd = O.flat_topk(elem, metric, dim, rows[c], q[i:i + 1], k)
        assert np.array_equal(ids[i], c[wi[0]]), (i, ids[i], c[wi[0]])
        assert np.array_equal(dists[i], wd[0])
    ix.close()


def test_topk_merge_of_shards_equals_unsharded():
    import torch
    elem, metric, rows, q = prepared("vector_l2_ops", "f32", 9000, 64, 130, 51)
    k, parts = 10, 3
    bounds = [0, 2500, 2500, 9000]        # an empty shard in the middle
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(q).to(dev)
    p_ids = torch.empty((parts, len(q), k), dtype=torch.int32, device=dev)
    p_d = torch.empty((parts, len(q), k), dtype=torch.float32, device=dev)
    handles = []
    for s in range(parts):
        lo, hi = bounds[s], bounds[s + 1]
        ix = P.HnswIndex("vector_l2_ops", 64, capacity=max(hi - lo, 1))
        ix.load_rows(rows[lo:hi])
        ix.flat_topk_device(d_q.data_ptr(), len(q), k, p_ids[s].data_ptr(), p_d[s].data_ptr(), id_base=lo)
        handles.append(ix)
    torch.cuda.synchronize()
    o_ids = torch.empty((len(q), k), dtype=torch.int32, device=dev)
    o_d = torch.empty((len(q), k), dtype=torch.float32, device=dev)
    handles[0].topk_merge_device(p_ids.data_ptr(), p_d.data_ptr(), parts, len(q), k, k, o_ids.data_ptr(), o_d.data_ptr())
    torch.cuda.synchronize()
    wi, wd = O.flat_topk(elem, metric, 64, rows, q, k, nthreads=THREADS)
    assert np.array_equal(o_ids.cpu().numpy(), wi) and np.array_equal(o_d.cpu().numpy(), wd)
    for ix in handles:
        ix.close()
