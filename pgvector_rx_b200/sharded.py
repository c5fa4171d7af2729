"""Multi-GPU plumbing for the two paths that shard, one process per GPU.

* HNSW scan: queries are independent, the graph is replicated on every GPU.  `query_shard` gives each
  rank a contiguous slice of the batch; there is no collective on the data path.
* Flat kNN / re-rank: the base set is split into contiguous row shards, every rank sees all queries,
  computes its local top-k with global ids (`id_base` = first row of its shard), the per-rank lists are
  all-gathered (NCCL over NVLink: nq * k * 8 bytes per rank) and merged by the K6 kernel on every rank.
* Index build does not shard (insert i depends on every insert before it): replicas only.

torch.distributed is used for the collective only.  The compute steps are the CUDA library's; the two
callables a `ShardedFlat` is built from exist so that the partitioning / gather-order logic can be run
under `gloo` on CPU tensors in the tests, where a checker stands in for the kernels.
"""


def shard_bounds(n, world):
    """Contiguous, balanced [lo, hi) ranges: the first n % world shards hold one extra item."""
    base, extra = divmod(int(n), int(world))
    bounds, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def query_shard(nq, rank, world):
    """Slice of a query batch owned by `rank` when queries shard against a replicated graph."""
    return shard_bounds(nq, world)[rank]


class ShardedFlat:
    """Exact kNN over a base set sharded by rows across the ranks of a process group.

    local_topk(queries, k) -> (ids, dists) tensors of shape [nq, k] with GLOBAL ids, ascending
    (distance, id), padded with -1 / +inf; merge(part_ids, part_dists, k) -> (ids, dists) takes the
    gathered [world, nq, k] tensors.  `for_index` wires both to the CUDA library."""

    def __init__(self, local_topk, merge, group=None):
        self.local_topk, self.merge, self.group = local_topk, merge, group

    @classmethod
    def for_index(cls, index, id_base, group=None, stream=0):
        """index: an HnswIndex whose loaded rows are this rank's shard; id_base: its first global row id."""
        import torch

        def local_topk(d_queries, k, out=None):
            nq = d_queries.shape[0]
            if out is None:
                ids = torch.empty((nq, k), dtype=torch.int32, device=d_queries.device)
                dists = torch.empty((nq, k), dtype=torch.float32, device=d_queries.device)
            else:
                ids, dists = out                       # the two halves of one packed send buffer
            index.flat_topk_device(d_queries.data_ptr(), nq, k, ids.data_ptr(), dists.data_ptr(), id_base=id_base,
                                   stream=stream or torch.cuda.current_stream().cuda_stream)
            return ids, dists

        def merge(part_ids, part_dists, k):
            world, nq, k_in = part_ids.shape
            ids = torch.empty((nq, k), dtype=torch.int32, device=part_ids.device)
            dists = torch.empty((nq, k), dtype=torch.float32, device=part_ids.device)
            s = stream or torch.cuda.current_stream().cuda_stream
            if part_ids.is_contiguous() and part_dists.is_contiguous():
                index.topk_merge_device(part_ids.data_ptr(), part_dists.data_ptr(), world, nq, k_in, k, ids.data_ptr(),
                                        dists.data_ptr(), stream=s)
            else:
                # views into a packed gather buffer: every part's list is contiguous, the parts are not adjacent
                index.topk_merge_parts_device([part_ids[w].data_ptr() for w in range(world)],
                                              [part_dists[w].data_ptr() for w in range(world)], nq, k_in, k,
                                              ids.data_ptr(), dists.data_ptr(), stream=s)
            return ids, dists

        return cls(local_topk, merge, group)

    def phase_times(self, queries, k, reps=5):
        """Per-phase device times of topk() (ms, CUDA events on the current stream, mean of `reps`, this rank): the local
        top-k over the shard, the all-gather(s), the K6 merge.  Diagnostic: tells which phase bounds strong scaling."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        acc = {"local_topk": 0.0, "all_gather": 0.0, "merge": 0.0}
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            if world > 1:
                send, ids, dists = self._packed(queries, k)
                self.local_topk(queries, k, out=(ids, dists))
            else:
                ids, dists = self.local_topk(queries, k)
            ev[1].record()
            if world > 1:
                part_ids, part_dists = self._gather_packed(send, world)
            else:
                part_ids, part_dists = ids.unsqueeze(0).contiguous(), dists.unsqueeze(0).contiguous()
            ev[2].record()
            self.merge(part_ids, part_dists, k)
            ev[3].record()
            torch.cuda.synchronize()
            acc["local_topk"] += ev[0].elapsed_time(ev[1])
            acc["all_gather"] += ev[1].elapsed_time(ev[2])
            acc["merge"] += ev[2].elapsed_time(ev[3])
        return {key: v / reps for key, v in acc.items()}

    def topk(self, queries, k):
        """All ranks call this with the same queries; every rank returns the merged global top-k."""
        import torch.distributed as dist
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            ids, dists = self.local_topk(queries, k)
            return self.merge(ids.unsqueeze(0).contiguous(), dists.unsqueeze(0).contiguous(), k)
        send, ids, dists = self._packed(queries, k)
        self.local_topk(queries, k, out=(ids, dists))
        # rank r's lists land in part[r]: the merge is order-independent for distinct ids, but keeping rank order makes
        # the result reproducible when shards hold identical (distance, id) pairs
        part_ids, part_dists = self._gather_packed(send, world)
        return self.merge(part_ids, part_dists, k)

    @staticmethod
    def _packed(queries, k):
        """One send buffer [2, nq, k] of 32-bit words: ids in the first half, distance bits in the second (a view)."""
        import torch
        send = torch.empty((2, queries.shape[0], k), dtype=torch.int32, device=queries.device)
        return send, send[0], send[1].view(torch.float32)

    def _gather_packed(self, send, world):
        """ONE all-gather of nq * k * 8 bytes per rank (at this size a collective is latency, not bandwidth); returns the
        [world, nq, k] id and distance views into the receive buffer."""
        import torch
        import torch.distributed as dist
        recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
        return recv[:, 0], recv[:, 1].view(torch.float32)


def gather_query_shards(local_ids, local_dists, nq, group=None):
    """Optional convenience for the query-sharded scan: collect every rank's result slice on all ranks
    (slices may differ by one row, so they are padded to the largest).  Not on the timed path of bench.py:
    results are disjoint slices that the caller normally consumes where they are."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bounds = shard_bounds(nq, world)
    width = max(hi - lo for lo, hi in bounds)
    k = local_ids.shape[1]

    def padded(t, fill):
        out = torch.full((width, k), fill, dtype=t.dtype, device=t.device)
        out[: t.shape[0]] = t
        return out

    all_ids = torch.empty((world, width, k), dtype=local_ids.dtype, device=local_ids.device)
    all_d = torch.empty((world, width, k), dtype=local_dists.dtype, device=local_dists.device)
    dist.all_gather_into_tensor(all_ids.view(-1, k), padded(local_ids, -1), group=group)
    dist.all_gather_into_tensor(all_d.view(-1, k), padded(local_dists, float("inf")), group=group)
    ids = torch.cat([all_ids[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)])
    dists = torch.cat([all_d[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)])
    return ids, dists
