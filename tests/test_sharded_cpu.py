"""World-size-2 `gloo` tests of the multi-GPU plumbing (pgvector_rx_b200/sharded.py) on CPU tensors:
row-shard bounds, global id offsets, all-gather order and the merge contract of the sharded flat path,
and the query-shard split of the replicated-graph scan.  The CUDA kernels are replaced by the oracle
here (this file tests the host-side logic; tests/test_flat_gpu.py and bench.py cover the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_ffi as O
from helpers import make_rows, oracle_index
from pgvector_rx_b200 import sharded as S


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 10_000, 1_000_003):
        for w in (1, 2, 3, 4, 8):
            b = S.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
            assert S.query_shard(n, w - 1, w) == b[-1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _numpy_merge(part_ids, part_dists, k):
    """Checker for the K6 contract: k smallest (distance, id) of the gathered lists, -1 entries skipped."""
    world, nq, k_in = part_ids.shape
    ids = np.full((nq, k), -1, np.int32)
    ds = np.full((nq, k), np.inf, np.float32)
    pi, pd = part_ids.numpy(), part_dists.numpy()
    for q in range(nq):
        cand = [(pd[r, q, j], pi[r, q, j]) for r in range(world) for j in range(k_in) if pi[r, q, j] >= 0]
        cand.sort()
        for j, (d, i) in enumerate(cand[:k]):
            ids[q, j], ds[q, j] = i, d
    return torch.from_numpy(ids), torch.from_numpy(ds)


def _worker(rank, world, port, n, dim, nq, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = make_rows("f32", n, dim, 61)
        q = make_rows("f32", nq, dim, 62)
        lo, hi = S.shard_bounds(n, world)[rank]

        def local_topk(queries, kk, out=None):
            ids, ds = O.flat_topk(O.ELEM_F32, O.L2SQ, dim, rows[lo:hi], queries.numpy(), kk)
            ids = np.where(ids >= 0, ids + lo, -1).astype(np.int32)       # global ids
            if out is not None:                  # the two halves of the packed send buffer (one all-gather per call)
                out[0].copy_(torch.from_numpy(ids))
                out[1].copy_(torch.from_numpy(ds))
                return out
            return torch.from_numpy(ids), torch.from_numpy(ds)

        flat = S.ShardedFlat(local_topk, _numpy_merge)
        ids, ds = flat.topk(torch.from_numpy(q), k)
        np.save(os.path.join(out_dir, "flat_ids_%d.npy" % rank), ids.numpy())
        np.save(os.path.join(out_dir, "flat_d_%d.npy" % rank), ds.numpy())

        # query-sharded scan against a replicated graph: each rank searches its slice only
        o = oracle_index("vector_l2_ops", dim, rows)
        qlo, qhi = S.query_shard(nq, rank, world)
        li, ld, _ = o.search_batch(q[qlo:qhi], ef_search=40, k=k)
        gi, gd = S.gather_query_shards(torch.from_numpy(li), torch.from_numpy(ld), nq)
        np.save(os.path.join(out_dir, "scan_ids_%d.npy" % rank), gi.numpy())
        np.save(os.path.join(out_dir, "scan_d_%d.npy" % rank), gd.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_flat_and_query_sharded_scan_gloo(tmp_path, world):
    n, dim, nq, k = 1501, 24, 41, 10          # odd sizes: uneven shards on both axes
    mp.spawn(_worker, args=(world, _free_port(), n, dim, nq, k, str(tmp_path)), nprocs=world, join=True)
    rows = make_rows("f32", n, dim, 61)
    q = make_rows("f32", nq, dim, 62)
    want_ids, want_d = O.flat_topk(O.ELEM_F32, O.L2SQ, dim, rows, q, k)
    o = oracle_index("vector_l2_ops", dim, rows)
    scan_ids, scan_d, _ = o.search_batch(q, ef_search=40, k=k)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("flat_ids_%d.npy" % r)), want_ids)
        assert np.array_equal(np.load(tmp_path / ("flat_d_%d.npy" % r)), want_d)
        assert np.array_equal(np.load(tmp_path / ("scan_ids_%d.npy" % r)), scan_ids)
        assert np.array_equal(np.load(tmp_path / ("scan_d_%d.npy" % r)), scan_d)


def test_single_process_path_needs_no_process_group():
    rows = make_rows("f32", 300, 8, 3)
    q = make_rows("f32", 5, 8, 4)

    def local_topk(queries, kk):
        ids, ds = O.flat_topk(O.ELEM_F32, O.L2SQ, 8, rows, queries.numpy(), kk)
        return torch.from_numpy(ids), torch.from_numpy(ds)

    ids, ds = S.ShardedFlat(local_topk, _numpy_merge).topk(torch.from_numpy(q), 4)
    wi, wd = O.flat_topk(O.ELEM_F32, O.L2SQ, 8, rows, q, 4)
    assert np.array_equal(ids.numpy(), wi) and np.array_equal(ds.numpy(), wd)
