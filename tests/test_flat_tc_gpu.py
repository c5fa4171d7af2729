"""GPU parity of the tensor-core flat path (K4b: 3xTF32 tcgen05 candidate generation + exact re-rank +
certification).  The final ids / distances must be the oracle's exact result bit for bit, like the CUDA-core
path's; the approximate scores must stay inside the bound the certification relies on."""
import os

import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P
from helpers import make_rows

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


@pytest.fixture
def tc_path():
    old = os.environ.get("PGVRX_FLAT_PATH")
    os.environ["PGVRX_FLAT_PATH"] = "tc"
    yield
    if old is None:
        del os.environ["PGVRX_FLAT_PATH"]
    else:
        os.environ["PGVRX_FLAT_PATH"] = old


def all_distances(metric, dim, rows, q):
    """nq x n matrix of the reference's exact distances (the oracle's sorted full scan, un-sorted again)."""
    ids, ds = O.flat_topk(O.ELEM_F32, metric, dim, rows, q, len(rows), nthreads=THREADS)
    out = np.empty((len(q), len(rows)), np.float32)
    for i in range(len(q)):
        out[i, ids[i]] = ds[i]
    return out


@pytest.mark.parametrize("opclass,metric", [("vector_l2_ops", O.L2SQ), ("vector_ip_ops", O.NEG_IP)])
@pytest.mark.parametrize("n,dim,nq", [(1024, 128, 64), (700, 96, 130), (300, 33, 5)])
def test_tc_scores_within_certified_bound(opclass, metric, n, dim, nq):
    rows, q = make_rows("f32", n, dim, 81), make_rows("f32", nq, dim, 82)
    ix = P.HnswIndex(opclass, dim, capacity=n)
    ix.load_rows(rows)
    scores, eps = ix.flat_tc_scores(q)
    ref = all_distances(metric, dim, rows, q)
    err = np.abs(scores.astype(np.float64) - ref.astype(np.float64))
    worst = (err.max(axis=1) / eps).max()
    assert worst <= 1.0, "3xTF32 error exceeds the bound: %.3g of eps" % worst
    # the split GEMM really is ~fp32 accurate: far better than a single tf32 product (2^-11 relative) could be
    scale = np.abs(ref).max()
    assert err.max() <= 1e-5 * max(scale, 1.0), (err.max(), scale)
    ix.close()


def tc_case(opclass, metric, n, dim, nq, k, seed=91, rows=None, q=None, max_uncert=0.05):
    rows = make_rows("f32", n, dim, seed) if rows is None else rows
    q = make_rows("f32", nq, dim, seed + 1) if q is None else q
    ix = P.HnswIndex(opclass, dim, capacity=len(rows))
    ix.load_rows(rows)
    ids, dists, st = ix.flat_topk(q, k=k, with_stats=True)
    assert st["path"] in (1, 2) and st["n_candidates"] > 0
    wi, wd = O.flat_topk(O.ELEM_F32, metric, dim, rows, q, k, nthreads=THREADS)
    bad = np.nonzero((ids != wi).any(axis=1))[0]
    assert len(bad) == 0, "%d queries differ, first %d: %s vs %s" % (len(bad), bad[0], ids[bad[0]], wi[bad[0]])
    assert np.array_equal(dists, wd)
    assert st["n_uncertified"] <= max_uncert * len(q) + 1, st
    ix.close()
    return st


@pytest.mark.parametrize("opclass,metric", [("vector_l2_ops", O.L2SQ), ("vector_ip_ops", O.NEG_IP)])
def test_tc_c2_shape(tc_path, opclass, metric):
    tc_case(opclass, metric, 30000, 128, 1000, 10)


@pytest.mark.parametrize("n,dim,nq,k", [(5000, 96, 129, 10), (777, 130, 64, 24), (128, 32, 128, 1), (129, 33, 127, 5),
                                        (4097, 64, 300, 16), (50, 8, 3, 10)])
def test_tc_ragged_shapes(tc_path, n, dim, nq, k):
    tc_case("vector_l2_ops", O.L2SQ, n, dim, nq, k, max_uncert=1.0 if n < 200 else 0.05)


def test_tc_ties_fall_back_to_exact(tc_path):
    # 64 rows at exactly the same distance from the query: the k-th neighbour ties with every candidate the kernel
    # kept and with rows it did not keep, so the bound cannot certify the query; it is re-run on the exact path
    # and the result is still the oracle's (ties by id)
    eye = np.eye(64, dtype=np.float32)
    rows = np.concatenate([eye, 2 * eye, 3 * eye])
    q = np.concatenate([np.zeros((1, 64), np.float32), make_rows("f32", 31, 64, 5)])
    st = tc_case("vector_l2_ops", O.L2SQ, None, 64, None, 10, rows=rows, q=q, max_uncert=1.0)
    assert st["n_uncertified"] > 0


def test_tc_duplicates_and_shard_ids(tc_path):
    import torch
    rows = make_rows("f32", 6000, 48, 7)
    rows[100:140] = rows[5]                    # 40 copies of one row: more duplicates than k
    q = np.concatenate([make_rows("f32", 200, 48, 8), rows[5:6]])
    ix = P.HnswIndex("vector_l2_ops", 48, capacity=len(rows))
    ix.load_rows(rows)
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(q).to(dev)
    ids = torch.empty((len(q), 10), dtype=torch.int32, device=dev)
    ds = torch.empty((len(q), 10), dtype=torch.float32, device=dev)
    st = P.FlatStats()
    ix.flat_topk_device(d_q.data_ptr(), len(q), 10, ids.data_ptr(), ds.data_ptr(), id_base=1_000_000, stats=st)
    assert st.path in (1, 2)
    wi, wd = O.flat_topk(O.ELEM_F32, O.L2SQ, 48, rows, q, 10, nthreads=THREADS)
    assert np.array_equal(ids.cpu().numpy(), wi + 1_000_000) and np.array_equal(ds.cpu().numpy(), wd)
    ix.close()


def test_auto_dispatch_picks_tensor_cores_for_large_batches_only():
    rows = make_rows("f32", 20000, 64, 3)
    ix = P.HnswIndex("vector_l2_ops", 64, capacity=len(rows))
    ix.load_rows(rows)
    assert ix.flat_topk(make_rows("f32", 600, 64, 4), k=10, with_stats=True)[2]["path"] in (1, 2)
    assert ix.flat_topk(make_rows("f32", 40, 64, 4), k=10, with_stats=True)[2]["path"] == 0      # small batch
    assert ix.flat_topk(make_rows("f32", 600, 64, 4), k=100, with_stats=True)[2]["path"] == 0    # k beyond the candidate lists
    ix.close()
    ix = P.HnswIndex("vector_l1_ops", 64, capacity=len(rows))
    ix.load_rows(rows)
    assert ix.flat_topk(make_rows("f32", 600, 64, 4), k=10, with_stats=True)[2]["path"] == 0      # L1 is not a contraction
    ix.close()


@pytest.mark.parametrize("knobs", [{}, {"PGVRX_TC_1CTA": "1"}, {"PGVRX_TC_NO_APPEND": "1"}, {"PGVRX_TC_PREPASS_SETS": "1"},
                                   {"PGVRX_TC_NO_PREPASS": "1"}])
@pytest.mark.parametrize("k", [10, 24])
def test_tc_prepass_and_append_forms_are_all_exact(tc_path, monkeypatch, knobs, k):
    # tables of 64+ tiles take the hint pre-pass (lane-local top lists of 8 keys for k <= 10, 16 beyond; or sets) and then
    # the append-only pass (or sets again); every combination must still return the oracle's exact result
    for name, value in knobs.items():
        monkeypatch.setenv(name, value)
    st = tc_case("vector_l2_ops", O.L2SQ, 40000, 64, 700, k, seed=95)
    assert st["n_uncertified"] == 0
    tc_case("vector_ip_ops", O.NEG_IP, 33000, 40, 520, k, seed=96)


def test_tc_append_overflow_falls_back_to_exact(tc_path):
    # 3000 copies of one row: a query next to it has thousands of rows under any threshold the sample can produce, its
    # list segments overflow, the query stays uncertified and is answered by the exact path -- identically
    rows = make_rows("f32", 40000, 32, 97)
    rows[1000:4000] = rows[7]
    q = np.concatenate([rows[7:8] + 1e-3, make_rows("f32", 600, 32, 98)]).astype(np.float32)
    st = tc_case("vector_l2_ops", O.L2SQ, None, 32, None, 10, rows=rows, q=q, max_uncert=1.0)
    assert st["n_uncertified"] >= 1
