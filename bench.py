#!/usr/bin/env python
"""bench.py -- QPS @ recall@10 of the batched HNSW index scan, plus build vectors/s, on the BASELINE.json configs.

    python bench.py [--config c2|c3|c4|c5] [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Default (the driver's contract): C2 = BASELINE.json configs[1], 1M x 128 fp32 vector_l2_ops, m=16,
ef_construction=64, ef_search=40, k=10, 10k-query batches.  A step = one pass of the hot path (get_scan_items for every
query of one 10k batch) on each GPU.  Queries shard across GPUs against a replicated graph (weak scaling: 10k queries per
GPU per step, no data-path collective).  Prints ONE JSON line on rank 0.  The C2 line also carries ef_search=100 (the
other half of configs[1]), an ef_search sweep on two datasets (`recall_curve`), the exact flat kNN, the GPU index build
(adjacency compared with the oracle's sequential build), inserts, and at N > 1 a strong-scaling arm.

--config c3|c4|c5 run the other BASELINE configs at their stated sizes through the same steps (GPU build, scan, parity
sample against the CPU oracle on the same graph, roofline from the kernel's own counters, CPU baseline, exact kNN for
recall), plus what the config names: C4 the relaxed-order iterative scan under a filter, C5 the exact re-rank and, at
N > 1, the row-sharded flat top-k with the NCCL merge.  --n / --nq scale any config down (tests, quick looks).

The C2 graph is the reference-exact graph built by the CPU oracle's sequential ambuild loop, cached under data_cache/
(tests/tools/build_graph_cache.py; ~50 minutes for 1M rows); the other configs' graphs are built on the GPU by
pgvrx_index_build, whose adjacency is asserted identical to the oracle's in tests/ and, for C2, in every run of this
file.  Vectors and queries are regenerated from seeds.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

_JSON_OUT = sys.stdout
METRIC_C2 = "QPS@recall@10 and build vectors/s, 1M×128 L2 HNSW, 1/2/4/8 B200"

# BASELINE.json configs[1..4] (SURVEY.md 8d fixes the seeds)
CONFIGS = {
    "c2": dict(name="C2", opclass="vector_l2_ops", n=1_000_000, dim=128, m=16, efc=64, ef=40, data="gaussian",
               data_seed=2, query_seed=3, level_seed=4, dtype="f32"),
    "c3": dict(name="C3", opclass="halfvec_cosine_ops", n=1_000_000, dim=768, m=32, efc=128, ef=40, data="gaussian_f16",
               data_seed=5, query_seed=6, level_seed=4, dtype="f16 rows, f32 accumulation"),
    "c4": dict(name="C4", opclass="vector_ip_ops", n=10_000_000, dim=96, m=16, efc=64, ef=40, data="gaussian",
               data_seed=7, query_seed=27, level_seed=4, dtype="f32", iterative=True, filters=(50, 500), max_scan_tuples=20000),
    "c5": dict(name="C5", opclass="bit_hamming_ops", n=5_000_000, dim=1024, m=16, efc=64, ef=40, data="clustered_bits",
               data_seed=8, query_seed=28, level_seed=4, dtype="u32 popcount", rerank=True),
}


def host_cores(a=None):
    """Host threads for the CPU legs: --threads, else the cores this process is allowed to run on."""
    if a is not None and getattr(a, "threads", 0) > 0:
        return a.threads
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=0, help="rows (0 = the config's size)")
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--m", type=int, default=0)
    ap.add_argument("--efc", type=int, default=0)
    ap.add_argument("--ef", type=int, default=0)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--cache", default=os.path.join(ROOT, "data_cache"))
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--recall-sample", type=int, default=200)
    ap.add_argument("--parity-sample", type=int, default=300, help="queries compared id for id (and counter for counter) with the oracle")
    ap.add_argument("--build-rows", type=int, default=-1,
                    help="C2: rows of the same dataset built on the GPU for the build vectors/s figure (0 = skip)")
    ap.add_argument("--insert-rows", type=int, default=-1,
                    help="C2: rows inserted into the finished index through aminsert's rules (-1 = 20000 when the whole dataset was built at N=1, else 0)")
    ap.add_argument("--vacuum-rows", type=int, default=0,
                    help="C2, after the insert sample: delete this many rows of the built index and VACUUM it on the GPU and in the "
                         "oracle, comparing the graphs (0 = skip; every deleted row makes ~2m live elements need a repair search)")
    ap.add_argument("--threads", type=int, default=0, help="host threads for the CPU legs (0 = the cores this process may use)")
    ap.add_argument("--no-flat", action="store_true", help="skip the exact flat-kNN section (recall then uses the oracle's brute force on --recall-sample queries)")
    ap.add_argument("--no-recall-curve", action="store_true", help="C2: skip the ef_search sweep (gaussian and low-intrinsic-dimension data)")
    ap.add_argument("--no-ef100", action="store_true", help="C2: skip the ef_search = 100 section")
    ap.add_argument("--no-gather-probe", action="store_true", help="skip the gather-ceiling probe")
    ap.add_argument("--iter-queries", type=int, default=10000, help="C4: queries per filtered iterative-scan batch (a quarter of it under the 1-in-500 filter)")
    ap.add_argument("--sweep", action="store_true", help="also time the scan under other staging geometries (tuning; rank 0, N = 1)")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    for key, attr in (("n", "n"), ("dim", "dim"), ("m", "m"), ("efc", "efc"), ("ef", "ef")):
        if getattr(a, attr) <= 0:
            setattr(a, attr, c[key])
    a.cfg = c
    a.full_size = (a.n, a.dim, a.m, a.efc) == (c["n"], c["dim"], c["m"], c["efc"])
    if a.build_rows < 0:
        # the build does not shard (replicas only): the full 1M-row build is measured at N = 1, larger N keep a 100k-row
        # sample of it on rank 0 so that the scaling runs stay short
        a.build_rows = a.n if int(os.environ.get("WORLD_SIZE", "1")) == 1 else min(a.n, 100_000)
    return a


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def raw_rows(a, n, seed, kind=None):
    """Seeded synthetic rows of a config, as the TABLE holds them (before the opclass normalises anything)."""
    from pgvector_rx_b200 import datasets as D
    kind = kind or a.cfg["data"]
    if kind == "gaussian":
        return D.gaussian_f32(n, a.dim, seed)
    if kind == "gaussian_f16":
        # fp32 N(0,1) -> halfvec (IEEE round-to-nearest-even); the cosine opclass normalises on the way in
        x = D.gaussian_f32(n, a.dim, seed)
        out = np.empty((n, a.dim), np.uint16)
        step = 1 << 16
        for s in range(0, n, step):
            out[s:s + step] = x[s:s + step].astype(np.float16).view(np.uint16)
        return out
    if kind == "clustered_bits":
        return D.clustered_bits(n, a.dim, seed)
    if kind == "lowdim":
        return D.low_dim_f32(n, a.dim, seed)
    raise ValueError(kind)


def workload_text(a, extra=""):
    c = a.cfg
    size = {1_000_000: "1M", 10_000_000: "10M", 5_000_000: "5M"}.get(a.n, str(a.n))
    unit = "bit" if c["opclass"].startswith("bit") else ("fp16" if c["opclass"].startswith("halfvec") else "fp32")
    return "%s: %sx%d %s %s m=%d ef_construction=%d ef_search=%d k=%d, %d-query batches%s" % (
        c["name"], size, a.dim, unit, c["opclass"], a.m, a.efc, a.ef, a.k, a.nq, extra)


def metric_name(a):
    if a.config == "c2":
        return METRIC_C2
    return "QPS@recall@10 and build vectors/s, " + workload_text(a).split(",")[0]


def graph_tag(a):
    if a.config == "c2":
        return "c2_1m" if (a.n, a.dim, a.m, a.efc) == (1_000_000, 128, 16, 64) else "g_%d_%d_%d_%d" % (a.n, a.dim, a.m, a.efc)
    return None


def load_cached_graph(a):
    tag = graph_tag(a)
    if not tag:
        return None
    d = os.path.join(a.cache, tag)
    meta_p = os.path.join(d, "meta.json")
    if not os.path.exists(meta_p):
        return None
    meta = json.load(open(meta_p))
    if (meta["n"], meta["dim"], meta["m"], meta["efc"], meta["data_seed"], meta["level_seed"]) != (a.n, a.dim, a.m, a.efc, 2, 4):
        return None
    g = {k: np.load(os.path.join(d, k + ".npy")) for k in ("level", "adj0", "upper_base", "adj_up")}
    g.update(rows=raw_rows(a, a.n, a.cfg["data_seed"]), entry=meta["entry"], n=meta["n"], source="cache:" + tag,
             build_vec_per_s_cpu=meta["n"] / meta["build_seconds"])
    return g


def gpu_build_graph(a, local, extra_capacity=0):
    """CREATE INDEX on the GPU (pgvrx_index_build).  Returns (index with the graph resident, graph arrays, stats)."""
    import pgvector_rx_b200 as P
    from pgvector_rx_b200 import datasets as D
    c = a.cfg
    rows = raw_rows(a, a.n, c["data_seed"])
    lv = D.hnsw_levels(a.n, a.m, c["level_seed"])
    ix = P.HnswIndex(c["opclass"], a.dim, m=a.m, ef_construction=a.efc, capacity=a.n + extra_capacity, device=local)
    t0 = time.perf_counter()
    st = ix.build(rows, lv)
    dt = time.perf_counter() - t0
    del rows
    g = ix.export_graph()
    g = dict(rows=g["rows"], level=g["level"], adj0=g["adj0"], upper_base=g["upper_base"], adj_up=g["adj_up"], entry=g["entry"],
             n=g["n"], source="gpu-build", build_vec_per_s_cpu=None)
    info = dict(gpu_vectors_per_s=g["n"] / dt, rows=int(g["n"]), seconds=dt, rounds=st["n_rounds"], searches=st["n_searched"],
                conflicts=st["n_conflicts"], big_retries=st["n_big_retries"], kernel_ms=st["kernel_ms"], gpu_ms=st["gpu_ms"],
                host_threads=st["host_threads"], commits_per_round=g["n"] / max(st["n_rounds"], 1),
                pairwise_on_gpu=st["n_pair_gpu"], host_distance_evals=st["n_host_dist"],
                search_counters=dict(n_dist=st["n_dist"], n_expand0=st["n_expand0"], n_expand_upper=st["n_expand_upper"]),
                note="pgvrx_index_build: speculative GPU search batches + host heuristic in the reference's order; exact "
                     "sequential semantics (adjacency identical to the oracle's: tests/test_build_gpu.py, and C2 at 1M in every "
                     "default run of this file)")
    return ix, g, info


def load_or_build_graph(a, local=0, want_index=False):
    """Returns (graph dict, index or None, gpu build info or None).  C2 prefers the oracle-built cache."""
    g = load_cached_graph(a)
    if g is not None:
        return g, None, None
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        log("no cached graph for %s: building %d rows on the GPU" % (a.config, a.n))
        ix, g, info = gpu_build_graph(a, local)
        if not want_index:
            ix.close()
            ix = None
        return g, ix, info
    log("no cached graph and no GPU: building %d rows with the CPU oracle (slow)" % a.n)
    import oracle_ffi as O
    import pgvector_rx_b200 as P
    from pgvector_rx_b200 import datasets as D
    elem, metric, norm = P.OPCLASSES[a.cfg["opclass"]]
    rows = raw_rows(a, a.n, a.cfg["data_seed"])
    if norm:
        rows = P.normalize_rows(elem, a.dim, rows)[1]
    o = O.Oracle(elem, metric, a.dim, m=a.m, ef_construction=a.efc, tie_mode=O.TIE_TOTAL_ORDER)
    t0 = time.time()
    o.build(rows, D.hnsw_levels(a.n, a.m, a.cfg["level_seed"]))
    dt = time.time() - t0
    g = o.export_graph()
    g.update(source="oracle-build", build_vec_per_s_cpu=a.n / dt)
    tag = graph_tag(a)
    if tag:
        try:
            d = os.path.join(a.cache, tag)
            os.makedirs(d, exist_ok=True)
            for k in ("level", "adj0", "upper_base", "adj_up"):
                np.save(os.path.join(d, k + ".npy"), g[k])
            json.dump(dict(n=int(g["n"]), dim=a.dim, m=a.m, efc=a.efc, data_seed=2, level_seed=4, entry=int(g["entry"]),
                           build_seconds=dt), open(os.path.join(d, "meta.json"), "w"))
        except OSError:
            pass
    return g, None, None


def query_batches(a, count, rank, prepare=None, kind=None, seed0=None):
    """`count` rotating query batches of this rank, in the form the scan takes them (normalised for cosine opclasses)."""
    s0 = a.cfg["query_seed"] if seed0 is None else seed0
    out = []
    for i in range(count):
        q = raw_rows(a, a.nq, s0 + 1000 * rank + 17 * i, kind)
        out.append(np.ascontiguousarray(prepare(q) if prepare else q))
    return out


def oracle_for(a, g):
    import oracle_ffi as O
    import pgvector_rx_b200 as P
    elem, metric, _ = P.OPCLASSES[a.cfg["opclass"]]
    o = O.Oracle(elem, metric, a.dim, m=a.m, ef_construction=a.efc, tie_mode=O.TIE_TOTAL_ORDER)
    o.import_graph(g)
    return o


def cpu_baseline(a, o, queries, ef, sample=0, budget_s=12.0):
    """The oracle (CPU restatement of the reference's scan) on all host cores: passes over the same 10k-query batch (or
    `sample` queries of it) until about budget_s seconds of wall time are used."""
    import oracle_ffi as O
    cores = host_cores(a)
    q = queries[:sample] if sample else queries
    o.search_batch(q[: 8 * cores], ef_search=ef, k=a.k, nthreads=cores)      # warm caches / threads / clocks
    done, t0, ids = 0, time.perf_counter(), None
    while True:
        r = o.search_batch(q, ef_search=ef, k=a.k, nthreads=cores)
        ids = r[0] if ids is None else ids
        done += len(q)
        dt = time.perf_counter() - t0
        if dt >= budget_s or sample:
            break
    return dict(value=done / dt, unit="queries/s", cores=cores, kind="port",
                flags="-O3 -march=native (built on this box)" if O.NATIVE else "-O2 (portable build)",
                sample="%d queries (%d passes over %d of the batch), ef_search=%d, k=%d, %d threads, %.1f s"
                       % (done, done // len(q), len(q), ef, a.k, cores, dt)), ids


def gpu_build_sample(a, local, n_insert=0):
    """C2: index build (CREATE INDEX) of the first --build-rows rows of the dataset through pgvrx_index_build: GPU-batched
    search_layer + host heuristic.  When an oracle graph of exactly that prefix is cached, the adjacency is compared with
    it (it must be identical)."""
    import pgvector_rx_b200 as P
    from pgvector_rx_b200 import datasets as D
    n = min(a.build_rows, a.n)
    X = raw_rows(a, a.n, a.cfg["data_seed"])
    X = X[:n] if n < a.n else X
    lv = D.hnsw_levels(a.n, a.m, a.cfg["level_seed"])[:n]
    ix = P.HnswIndex(a.cfg["opclass"], a.dim, m=a.m, ef_construction=a.efc, capacity=n + n_insert, device=local)
    t0 = time.perf_counter()
    st = ix.build(X, lv)
    dt = time.perf_counter() - t0
    out = dict(gpu_vectors_per_s=n / dt, rows=n, seconds=dt, rounds=st["n_rounds"], searches=st["n_searched"],
               conflicts=st["n_conflicts"], kernel_ms=st["kernel_ms"], gpu_ms=st["gpu_ms"], host_threads=st["host_threads"],
               commits_per_round=n / max(st["n_rounds"], 1), pairwise_on_gpu=st["n_pair_gpu"], host_distance_evals=st["n_host_dist"],
               note="%s rows of the C2 dataset (same seeds and levels as the oracle's build); exact sequential semantics, "
                    "see DESIGN.md" % ("all %d" % n if n == a.n else "first %d" % n))
    tag = "c2_1m" if (n, a.dim, a.m, a.efc) == (1_000_000, 128, 16, 64) else "g_%d_%d_%d_%d" % (n, a.dim, a.m, a.efc)
    d = os.path.join(a.cache, tag)
    if a.config == "c2" and os.path.exists(os.path.join(d, "meta.json")):
        meta = json.load(open(os.path.join(d, "meta.json")))
        g = ix.export_graph()
        same = g["entry"] == meta["entry"] and all(np.array_equal(g[k], np.load(os.path.join(d, k + ".npy")))
                                                   for k in ("level", "adj0", "upper_base", "adj_up"))
        out["adjacency_identical_to_oracle"] = bool(same)
        out["cpu_oracle_vectors_per_s_same_rows"] = meta["n"] / meta["build_seconds"]
    if n_insert > 0:
        return out, ix
    ix.close()
    return out, None


def insert_sample(a, ix, o, n_insert, n_check=1500):
    """INSERT into the finished index (aminsert, src/index/insert.rs): n_insert new rows through pgvrx_index_insert.
    The first n_check of them also go through the oracle's aminsert on the same graph and the two graphs are compared
    (every list of every layer); the rest measure throughput."""
    from pgvector_rx_b200 import datasets as D
    X = D.gaussian_f32(n_insert, a.dim, 12)
    lv = D.hnsw_levels(n_insert, a.m, 13)
    n_check = min(n_check, n_insert)
    n0 = ix.graph_info()["n"]
    t0 = time.perf_counter()
    o.insert(X[:n_check], lv[:n_check], base_tid=n0)
    dt_cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    st1 = ix.insert(X[:n_check], lv[:n_check], heap_tids=np.arange(n0, n0 + n_check, dtype=np.uint64))
    dt1 = time.perf_counter() - t0
    g, w = ix.export_graph(), o.export_graph()
    same = g["n"] == w["n"] and g["entry"] == w["entry"] and all(np.array_equal(g[k], w[k]) for k in ("level", "adj0", "upper_base", "adj_up"))
    del g, w
    t0 = time.perf_counter()
    st2 = ix.insert(X[n_check:], lv[n_check:], heap_tids=np.arange(n0 + n_check, n0 + n_insert, dtype=np.uint64))
    dt2 = time.perf_counter() - t0
    return dict(gpu_vectors_per_s=n_insert / (dt1 + dt2), rows=n_insert, into_rows=int(n0), seconds=dt1 + dt2,
                rounds=st1["n_rounds"] + st2["n_rounds"], conflicts=st1["n_conflicts"] + st2["n_conflicts"],
                host_distance_evals=st1["n_host_dist"] + st2["n_host_dist"],
                cpu_oracle_vectors_per_s=n_check / dt_cpu, cpu_sample_rows=n_check,
                graph_identical_to_oracle=bool(same),
                note="aminsert rules (nearest-lm lists, one-slot back-link replacement); the oracle inserts the first %d rows "
                     "single-threaded on the same graph and the resulting graphs are compared" % n_check)


def vacuum_sample(a, ix, o, n_dead):
    """VACUUM (ambulkdelete, src/index/vacuum.rs) after deleting n_dead rows spread over the index: pgvrx_index_vacuum on
    the GPU and the oracle's bulk_delete on the same graph, graphs and flags compared.  Off by default (--vacuum-rows)."""
    n = ix.graph_info()["n"]
    dead = np.random.default_rng(21).permutation(n)[:n_dead].astype(np.int64)
    t0 = time.perf_counter()
    o.bulk_delete(dead)
    dt_cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    st = ix.vacuum(dead.astype(np.uint64))
    dt = time.perf_counter() - t0
    g, w = ix.export_graph(), o.export_graph()
    same = g["entry"] == w["entry"] and all(np.array_equal(g[k], w[k]) for k in ("adj0", "adj_up")) and \
        np.array_equal(ix.element_flags(), o.element_flags())
    return dict(rows_deleted=int(n_dead), elements_repaired=int(st["n_elements_repaired"]), seconds=dt,
                repairs_per_s=st["n_elements_repaired"] / dt, rounds=int(st["n_rounds"]), conflicts=int(st["n_conflicts"]),
                kernel_ms=st["kernel_ms"], cpu_oracle_seconds=dt_cpu, cpu_oracle_repairs_per_s=st["n_elements_repaired"] / dt_cpu,
                graph_identical_to_oracle=bool(same))


def recall_at_k(a, g, queries, got_ids, sample):
    """recall@k against the oracle's brute force on a sample of the batch (CPU; used by the reference arm)."""
    import oracle_ffi as O
    import pgvector_rx_b200 as P
    elem, metric, _ = P.OPCLASSES[a.cfg["opclass"]]
    s = min(sample, len(queries))
    gt, _ = O.flat_topk(elem, metric, a.dim, g["rows"], queries[:s], a.k, nthreads=host_cores(a))
    hit = sum(len(set(got_ids[i].tolist()) & set(gt[i].tolist())) for i in range(s))
    return hit / float(s * a.k)


def recall_vs(gt_ids, got_ids):
    n, k = got_ids.shape
    hit = sum(len((set(got_ids[i].tolist()) - {-1}) & set(gt_ids[i].tolist())) for i in range(n))
    return hit / float(n * k)


def tf32_peak_probe(dev):
    """cuBLAS TF32 GEMM rate through torch, used for nothing but the flat path's roofline denominator
    (MEASURED_PEAKS.json has no tf32 figure); best of 10 like the driver's bf16 number."""
    import torch
    try:
        old_flag = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        ma = torch.randn((8192, 8192), device=dev)
        mb = torch.randn((8192, 8192), device=dev)
        for _ in range(3):
            mc = ma @ mb
        torch.cuda.synchronize()
        best = None
        for _ in range(10):
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            mc = ma @ mb
            p1.record()
            torch.cuda.synchronize()
            tms = p0.elapsed_time(p1)
            best = tms if best is None else min(best, tms)
        torch.backends.cuda.matmul.allow_tf32 = old_flag
        del ma, mb, mc
        return 2.0 * 8192 ** 3 / (best / 1e3) / 1e12
    except Exception as exc:                                    # the probe is optional
        log("tf32 peak probe failed: %r" % (exc,))
        return None


def flat_section(a, ix, g, batch, rank, world, local, dev, tstream):
    """Exact flat kNN: ground truth for recall and its own throughput line.

    N = 1: the whole table on one GPU (fp32 L2 / IP: 3xTF32 tcgen05 candidates + exact re-rank, certified; every other
    opclass: sequential f32 chains / popcounts on the CUDA cores).
    N > 1: the base set is sharded by rows, every rank sees the same query batch, local top-k lists are all-gathered over
    NCCL and merged by the K6 kernel on every rank (strong scaling of one batch)."""
    import torch
    import torch.distributed as dist
    import pgvector_rx_b200 as P
    from pgvector_rx_b200 import sharded as S
    stream = tstream.cuda_stream
    d_q = torch.from_numpy(batch.view(np.uint8) if batch.dtype != np.float32 else batch).to(dev)
    nq, k = len(batch), a.k
    st = P.FlatStats()
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    ds = torch.empty((nq, k), dtype=torch.float32, device=dev)
    phases = None
    if world == 1:
        fx = ix
        lo = 0
        run = lambda: fx.flat_topk_device(d_q.data_ptr(), nq, k, ids.data_ptr(), ds.data_ptr(), stream=stream)
        result = lambda: ids
    else:
        lo, hi = S.shard_bounds(a.n, world)[rank]
        fx = P.HnswIndex(a.cfg["opclass"], a.dim, m=a.m, ef_construction=a.efc, capacity=hi - lo, device=local)
        fx.load_rows_stored(g["rows"][lo:hi])
        sf = S.ShardedFlat.for_index(fx, lo, stream=stream)
        holder = {}

        def run():
            holder["r"] = sf.topk(d_q, k)
        result = lambda: holder["r"][0]
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 10 if a.n * a.dim <= 200_000_000 else 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = sf.phase_times(d_q, k)
    ms = float(t.item())
    fx.flat_topk_device(d_q.data_ptr(), nq, k, ids.data_ptr(), ds.data_ptr(), id_base=lo, stats=st)
    tc = st.path >= 1
    out = dict(value=nq / (ms / 1e3), unit="queries/s", ms_per_batch=ms, k=k, rows=a.n, queries=nq, n_gpus=world,
               path="3xTF32 tcgen05 candidates + exact re-rank (certified)" if tc else
                    ("popcount chains (CUDA cores)" if a.cfg["opclass"].startswith("bit") else "sequential f32 chains (CUDA cores)"),
               candidates_per_query=int(st.n_candidates // max(nq, 1)), uncertified=int(st.n_uncertified),
               parallelism="single table" if world == 1 else "base rows sharded x%d, NCCL all-gather of %d B/rank + K6 merge" % (world, nq * k * 8))
    if phases:
        out["phase_ms"] = phases
    if tc:
        useful = 2.0 * nq * a.n * a.dim
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        tf32_est = peaks.get("bf16_tflops", 2250.0 * 0.73) / 2.0
        tf32_meas = tf32_peak_probe(dev)
        tf32_peak = tf32_meas if tf32_meas else tf32_est
        issued = useful * 3 / world
        out["roofline"] = dict(bound="tensor", achieved=issued / (ms / 1e3) / 1e12, peak=tf32_peak, unit="TFLOP/s",
                               frac=issued / (ms / 1e3) / 1e12 / tf32_peak, traffic=None,
                               peak_source="torch.matmul TF32 8192^3, best of 10, this run" if tf32_meas else "measured dense bf16 / 2",
                               peak_bf16_half=tf32_est, frac_of_nominal_1125=issued / (ms / 1e3) / 1e12 / 1125.0,
                               note="issued tf32 flops per GPU (3 products x 2*Q*N*D / n_gpus) / CUDA-event time of the whole call "
                                    "(operand split, GEMM + top-k, re-rank, certification%s)" % (", all-gather, merge" if world > 1 else ""))
    else:
        # CUDA-core path: the bound is instruction issue (FP32 chains or POPC); report the table bytes
        # streamed per second next to the HBM peak for scale (every 32-query tile streams the table, mostly from L2)
        peak, peak_src = measured_peak()
        rb = fx.L.pgvrx_row_bytes(fx.elem, a.dim)
        pairs = float(nq) * a.n / world
        out["roofline"] = dict(bound="issue", achieved=pairs / (ms / 1e3) / 1e9, peak=None, unit="G pairs/s", frac=None, traffic=None,
                               table_bytes_streamed_per_s_GB=pairs / 32.0 * rb / (ms / 1e3) / 1e9, hbm_peak_GBps=peak,
                               note="(query, row) distance evaluations per second per GPU; one thread owns a row and 32 query chains")
    final_ids = result().cpu().numpy()
    if world > 1:
        fx.close()
    return out, final_ids


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.rows, self.proc, self.gpu = [], None, gpu
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """Start of the region whose samples are reported (the sampler itself starts before warm-up so that it is
        already streaming when the short timed region begins)."""
        self.marks.append(time.perf_counter())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        t0 = self.marks[0] if self.marks else 0.0
        rows = [r for t, r in self.rows if t >= t0] or [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm), samples_before_timed_region=len(self.rows) - len(rows),
                    region="device-timed steps + both end-to-end loops (all under load)")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key="scan_kernel"):
    """dram bytes per launch of the scan kernel from the committed ncu capture (a constant from profiles/, not measured
    by this run: ncu cannot run inside a timed bench); the file names the capture it came from."""
    p = os.path.join(ROOT, "profiles", kernel_key + "_traffic.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j.get("dram_bytes_per_launch"), j.get("source")
    return None, None


# ---------------------------------------------------------------------------------------------
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_ffi as O
    g, ixb, _ = load_or_build_graph(a)
    if ixb is not None:
        ixb.close()
    import pgvector_rx_b200 as P
    elem, metric, norm = P.OPCLASSES[a.cfg["opclass"]]
    prep = (lambda q: P.normalize_rows(elem, a.dim, q)[1]) if norm else None
    batches = query_batches(a, 1, 0, prep)
    o = oracle_for(a, g)
    cores = host_cores(a)
    per_step = a.nq if cores >= 8 else min(a.nq, 400 * cores)   # the whole 10k batch per step unless the box is tiny
    q = batches[0][:per_step]
    for _ in range(max(a.warmup, 1)):
        o.search_batch(q[: 8 * cores], ef_search=a.ef, k=a.k, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ids, _, _ = o.search_batch(q, ef_search=a.ef, k=a.k, nthreads=cores)
    dt = time.perf_counter() - t0
    qps = per_step * a.steps / dt
    rec = recall_at_k(a, g, q, ids, a.recall_sample)
    sample = "%d queries/step of the 10k batch, %d threads, %s" % (per_step, cores, "-O3 -march=native" if O.NATIVE else "-O2 portable build")
    out = dict(impl="reference", metric=metric_name(a), value=qps, unit="queries/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
               ms_per_step=dt / a.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=a.cfg["dtype"],
               data="synthetic", recall_at_10=rec,
               config=dict(workload=workload_text(a), n=a.n, dim=a.dim, queries_per_step_per_gpu=per_step,
                           parallelism="host threads x%d" % cores, graph=g["source"],
                           note="CPU restatement of pgvector-rx's scan on flat arrays (the Rust/pgrx reference cannot be built here)"),
               build=dict(cpu_oracle_vectors_per_s=g.get("build_vec_per_s_cpu"),
                          note="sequential ambuild restatement, one core (the reference's build is single-threaded per backend), "
                               "measured when the cached graph was built"),
               cpu_baseline=dict(value=qps, unit="queries/s", cores=cores, kind="port", sample=sample),
               e2e=dict(value=qps, unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), file=_JSON_OUT, flush=True)


class ScanRunner:
    """Device-timed and end-to-end scan loops on one index: two batches in flight on alternating streams (each with its own
    result buffers; the library keeps one scratch set per stream), so the ragged tail of one persistent launch -- warps
    finishing their last query at different times -- is filled by the next batch instead of idling the SMs."""

    def __init__(self, a, ix, batches, dev, world):
        import torch
        self.torch, self.a, self.ix, self.world, self.dev = torch, a, ix, world, dev
        self.nb, self.n_fly = len(batches), 2
        self.collective = True
        self.batches = batches
        as_u8 = lambda b: b if b.dtype == np.float32 else b.view(np.uint8)
        self.d_q = [torch.from_numpy(as_u8(b)).to(dev) for b in batches]
        self.tstreams = [torch.cuda.Stream(device=dev) for _ in range(self.n_fly)]
        self.pin_q = [torch.from_numpy(as_u8(b)).pin_memory() for b in batches]
        self.outs = {}
        self.pin_outs = {}

    def _bufs(self, k, pinned):
        torch, a = self.torch, self.a
        store = self.pin_outs if pinned else self.outs
        if k not in store:
            mk = (lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()) if pinned else \
                (lambda shape, dt: torch.empty(shape, dtype=dt, device=self.dev))
            store[k] = [(mk((a.nq, k), torch.int32), mk((a.nq, k), torch.float32), mk((a.nq,), torch.int32)) for _ in range(self.n_fly)]
        return store[k]

    def barrier(self):
        # `collective` is switched off for the sections only rank 0 runs (ef_search sweeps, CPU arms): a process-group
        # barrier there would wait for ranks that are not coming
        if self.world > 1 and self.collective:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def step(self, i, ef, k, pinned=False):
        o = self._bufs(k, pinned)[i % self.n_fly]
        q = (self.pin_q if pinned else self.d_q)[i % self.nb]
        self.ix.search_device(q.data_ptr(), self.a.nq, k, ef, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                              stream=self.tstreams[i % self.n_fly].cuda_stream)

    def device_timed(self, ef, k, steps, warmup):
        """CUDA events on the launch streams around `steps` steps; max over ranks.  Returns ms for all steps."""
        torch = self.torch
        for i in range(warmup):
            self.step(i, ef, k)
        self.barrier()
        t0s = self.tstreams[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for t in self.tstreams[1:]:
            t.wait_stream(t0s)
        e0.record(t0s)
        for t in self.tstreams[1:]:                    # no stream starts before the start event
            ev = torch.cuda.Event()
            ev.record(t0s)
            t.wait_event(ev)
        for i in range(steps):
            self.step(i, ef, k)
        for t in self.tstreams[1:]:                    # the timed region ends when every stream has drained
            ev = torch.cuda.Event()
            ev.record(t)
            t0s.wait_event(ev)
        e1.record(t0s)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1 and self.collective:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def wall_timed(self, step_fn, steps, warmup):
        torch = self.torch
        for i in range(max(3, warmup // 2)):
            step_fn(i)
        self.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            step_fn(i)
        self.barrier()
        dt = time.perf_counter() - t0
        if self.world > 1 and self.collective:
            import torch.distributed as dist
            tt = torch.tensor([dt], dtype=torch.float64, device=self.dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt

    def sync_step(self, i, ef, k):
        import ctypes as C
        o = self._bufs(k, True)[0]
        rc = self.ix.L.pgvrx_search_batch(self.ix.h, C.c_void_p(self.pin_q[i % self.nb].data_ptr()), self.a.nq, ef, k,
                                          C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()), C.c_void_p(o[2].data_ptr()), None)
        self.ix._check(rc)


def parity_sample(a, ix, o, queries, ef, k, ns):
    """GPU scan vs the oracle's scan of the same graph on ns queries: ids, distance bits and the work counters."""
    ns = min(ns, len(queries))
    if ns <= 0:
        return None
    got = ix.search(queries[:ns], k=k, ef_search=ef, with_stats=True, prepared=True)
    want = o.search_batch(queries[:ns], ef_search=ef, k=k, nthreads=host_cores(a), with_counters=True)
    ctr = all(got[3][key] == want[3][key] for key in ("n_dist", "n_expand0", "n_expand_upper", "n_visited"))
    return dict(queries=ns, ef_search=ef, ids_identical=bool(np.array_equal(got[0], want[0])),
                distances_bit_identical=bool(np.array_equal(got[1], want[1])), counters_identical=bool(ctr),
                n_retried=int(got[3]["n_retried"]))


def scan_roofline(a, ix, st, step_s, kernel):
    rb = ix.L.pgvrx_row_bytes(ix.elem, a.dim)
    row_stride = (rb + 15) // 16 * 16
    bytes_step = st.n_dist * row_stride + st.n_expand0 * 2 * a.m * 4 + st.n_expand_upper * a.m * 4
    peak, peak_src = measured_peak()
    achieved = bytes_step / step_s / 1e9
    traffic, traffic_src = ncu_traffic() if a.config == "c2" and a.full_size and a.ef == 40 else (None, None)
    return dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, frac_of_nominal_8000=achieved / 8000.0,
                traffic=traffic, traffic_source=traffic_src, peak_source=peak_src, kernel=kernel,
                bytes_per_launch=int(bytes_step), bytes_per_query=bytes_step / a.nq,
                counters=dict(n_dist=int(st.n_dist), n_expand0=int(st.n_expand0), n_expand_upper=int(st.n_expand_upper),
                              n_retried=int(st.n_retried)),
                note="achieved = (n_dist*%d + n_expand0*%d + n_expand_upper*%d bytes per step) / CUDA-event step time, rank 0 counters"
                     % (row_stride, 2 * a.m * 4, a.m * 4))


def staging_sweep(a, runner):
    """Device-resident QPS of the scan under other staging geometries (environment knobs of capi.cu's launch
    configuration), same batches, same two streams.  A tuning aid: which geometry is the default is decided from these."""
    knobs = [dict(), dict(PGVRX_STAGE_ROWS="32", PGVRX_SPEC_GATHER="1"), dict(PGVRX_STAGE_ROWS="32", PGVRX_SPEC_GATHER="0"),
             dict(PGVRX_STAGE_ROWS="16", PGVRX_SPEC_GATHER="1"), dict(PGVRX_STAGE_ROWS="16", PGVRX_SPEC_GATHER="0"),
             dict(PGVRX_STAGE_ROWS="16", PGVRX_SPEC_GATHER="0", PGVRX_SCAN_WARPS="2"), dict(PGVRX_STAGE_ROWS="11"), dict(PGVRX_STAGE_ROWS="8"),
             dict(PGVRX_CHUNK_BYTES="128"), dict(PGVRX_CHUNK_BYTES="256"), dict(PGVRX_CHUNK_BYTES="256", PGVRX_SPEC_GATHER="1"),
             dict(PGVRX_CHUNK_BYTES="384"), dict(PGVRX_CHUNK_BYTES="512"), dict(PGVRX_CHUNK_BYTES="512", PGVRX_SPEC_GATHER="1"),
             dict(PGVRX_CHUNK_BYTES="768"), dict(PGVRX_NO_CHUNK="1", PGVRX_STAGE_KB="17"), dict()]
    out = []
    steps = max(6, a.steps // 8)
    for kn in knobs:
        for k_, v in kn.items():
            os.environ[k_] = v
        try:
            ms = runner.device_timed(a.ef, a.k, steps, 3)
            out.append(dict(knobs=kn, qps=a.nq * steps / (ms / 1e3), ms_per_step=ms / steps))
        except Exception as exc:
            out.append(dict(knobs=kn, error=repr(exc)[:200]))
        for k_ in kn:
            del os.environ[k_]
    return out


def gather_probe_section(ix):
    """Measured ceiling for the scan's gathers (pgvrx_gather_probe): random row gathers with the scan's own TMA mechanism
    and geometry, ids from a random walk over the graph (hub reuse as in real searches) and uniform over the table."""
    out = {}
    try:
        for name, mode in (("graph_walk", 1), ("uniform", 0)):
            r = ix.gather_probe(mode=mode, steps=192, depth=1)
            r2 = ix.gather_probe(mode=mode, steps=192, depth=2, rows_per_step=16)
            out[name] = dict(gb_per_s=r["gb_per_s"], warps_per_sm=r["warps_per_sm"], rows_per_phase=32, phases_in_flight=1,
                             gb_per_s_two_half_phases=r2["gb_per_s"], bytes_per_launch=r["bytes_per_launch"], ms=r["ms_per_launch"])
    except Exception as exc:
        out["error"] = repr(exc)
    return out


def iterative_section(a, ix, o, queries, filt, nq_it, limit=10):
    """C4: hnsw.iterative_scan = relaxed_order under a filter the index cannot see (WHERE id % c = 0, tests/t/044...pl:13):
    the consumer keeps matching elements and stops a query after `limit` hits; rounds run for all unfinished queries at
    once (pgvrx_scan_next).  Streams of a few queries are compared with the oracle's amgettuple restatement."""
    import oracle_ffi as O
    import pgvector_rx_b200 as P
    nq_it = min(nq_it, len(queries))
    q = queries[:nq_it]
    mst = a.cfg.get("max_scan_tuples", 20000)
    t0 = time.perf_counter()
    sc = ix.scan_begin(q, ef_search=a.ef, iterative_scan=P.ITER_RELAXED, max_scan_tuples=mst, prepared=True)
    hit_ids = np.full((nq_it, limit), -1, np.int64)    # the consumer's rows, in the order it received them
    seen = np.zeros(nq_it, np.int64)                   # elements each query's consumer has looked at
    n_hits = np.zeros(nq_it, np.int64)
    done = np.zeros(nq_it, bool)
    rounds = 0
    emitted = 0
    # refills per call (pgvrx_scan_next_rounds): a filter that keeps 1 row in c needs about limit * c / ef_search refills;
    # asking for a quarter of that at a time bounds the work wasted on a query that finishes early to ~25 % of a scan
    rpc = int(max(1, min(16, limit * filt // (a.ef * 4))))
    col = np.arange(rpc * a.ef)[None, :]
    calls = 0
    while True:
        want = ~done & (n_hits < limit)
        if not want.any():
            break
        ii, dd, cc, dn = sc.next(want.astype(np.uint8), rounds=rpc)
        rounds += rpc
        calls += 1
        # the executor's filter, for all queries of the call at once: matching elements in emission order (batch after
        # batch), as many as the query still lacks
        live = (np.arange(a.ef)[None, None, :] < cc[:, :, None]) & want[:, None, None]
        ids_flat = ii.reshape(nq_it, rpc * a.ef)
        live = live.reshape(nq_it, rpc * a.ef)
        emitted += int(live.sum())
        match = live & (ids_flat % filt == 0)
        rank = np.cumsum(match, axis=1) - 1 + n_hits[:, None]          # slot each match would take
        take = match & (rank < limit)
        # elements the consumer looked at: everything up to its last taken row when that completed the limit, else all
        full = (n_hits + take.sum(axis=1)) >= limit
        last = np.where(take.any(axis=1), take.shape[1] - 1 - np.argmax(take[:, ::-1], axis=1), -1)
        looked = np.where(full, np.cumsum(live, axis=1)[np.arange(nq_it), np.maximum(last, 0)] * (last >= 0), live.sum(axis=1))
        seen += looked
        r_, c_ = np.nonzero(take)
        hit_ids[r_, rank[r_, c_]] = ids_flat[r_, c_]
        n_hits += take.sum(axis=1)
        done |= dn.astype(bool)
    dt = time.perf_counter() - t0
    hits = [[int(v) for v in hit_ids[i, :n_hits[i]]] for i in range(nq_it)]
    tot = dict(sc.totals)
    sc.close()
    # the oracle's amgettuple stream for a sample of the queries, as far as the GPU consumer had to look (elements hold one
    # heap TID each here): parity of the filtered rows, and the CPU arm of this section -- one query per host thread
    # (one backend = one thread in the reference), wall clock
    from concurrent.futures import ThreadPoolExecutor
    cores = host_cores(a)
    n_chk = min(max(4 * cores, 48), nq_it)

    def cpu_stream(i):
        _, eids, _, _ = o.scan_stream(q[i], ef_search=a.ef, iterative=O.ITER_RELAXED, max_scan_tuples=mst, max_emit=int(seen[i]))
        return [int(e) for e in eids if e % filt == 0][:limit]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        cpu_hits = list(ex.map(cpu_stream, range(n_chk)))
    cpu_dt = time.perf_counter() - t0
    ok = all(cpu_hits[i] == hits[i] for i in range(n_chk))
    rb = ix.L.pgvrx_row_bytes(ix.elem, a.dim)
    gathered = tot["n_dist"] * ((rb + 15) // 16 * 16) + tot["n_expand0"] * 2 * a.m * 4 + tot["n_expand_upper"] * a.m * 4
    return dict(filter="id %% %d == 0" % filt, limit=limit, queries=nq_it, value=nq_it / dt, unit="queries/s (end to end, host buffers)",
                seconds=dt, rounds=rounds, refills_per_call=rpc, calls=calls, kernel_ms=tot["kernel_ms"], elements_emitted_per_query=emitted / nq_it,
                queries_with_all_hits=int((n_hits >= limit).sum()), max_scan_tuples=mst,
                gathered_GBps_in_kernels=gathered / (tot["kernel_ms"] / 1e3) / 1e9 if tot["kernel_ms"] else None,
                streams_match_oracle=bool(ok), streams_checked=n_chk,
                cpu_baseline=dict(value=n_chk / cpu_dt, unit="queries/s", cores=cores, kind="port",
                                  sample="%d queries, the same filter, limit and stream length as the GPU consumer, one query per host thread, %.1f s"
                                         % (n_chk, cpu_dt)))


def rerank_section(a, ix, o, queries, gt_ids, ns=2000, k_prime=100):
    """C5: HNSW candidates (ef_search = k' = 100) re-ranked exactly (pgvrx_rerank) against the flat ground truth."""
    ns = min(ns, len(queries))
    q = queries[:ns]
    ef = max(a.ef, k_prime)
    ix.search(q[:64], k=k_prime, ef_search=ef, prepared=True)        # scratch for this ef / k is allocated on first use
    ix.rerank(q[:64], np.zeros((64, k_prime), np.int32), k=a.k, prepared=True)
    t0 = time.perf_counter()
    ids, dists, counts = ix.search(q, k=k_prime, ef_search=ef, prepared=True)
    t_scan = time.perf_counter() - t0
    t0 = time.perf_counter()
    ri, rd = ix.rerank(q, ids, k=a.k, prepared=True)
    t_rr = time.perf_counter() - t0
    out = dict(queries=ns, k_prime=k_prime, ef_search=ef, scan_seconds=t_scan, rerank_seconds=t_rr,
               value=ns / (t_scan + t_rr), unit="queries/s (scan + re-rank, host buffers)",
               rerank_equals_scan_head=bool(np.array_equal(ri, ids[:, :a.k])))
    if gt_ids is not None:
        out["recall_at_10_after_rerank"] = recall_vs(gt_ids[:ns], ri)
    return out


def ef_sweep(a, ix, o, runner, gt_ids, batch, efs, label, cpu_budget_s=2.5):
    """QPS and recall@10 per ef_search for both arms on one graph: GPU device-resident QPS (CUDA events), CPU oracle on a
    sample of the same batch on all cores."""
    rows = []
    for ef in efs:
        k = a.k
        steps = max(4, min(20, int(40 * 40 / ef)))
        ms = runner.device_timed(ef, k, steps, 3)
        qps = a.nq * steps / (ms / 1e3)
        ids, _, _ = ix.search(batch, k=k, ef_search=ef, prepared=True)
        ns = int(max(64, min(len(batch), 2000 * 40 / ef)))
        cores = host_cores(a)
        o.search_batch(batch[: 4 * cores], ef_search=ef, k=k, nthreads=cores)
        t0 = time.perf_counter()
        done = 0
        cpu_ids = None
        while True:
            r = o.search_batch(batch[:ns], ef_search=ef, k=k, nthreads=cores)
            cpu_ids = r[0]
            done += ns
            if time.perf_counter() - t0 >= cpu_budget_s:
                break
        cpu_qps = done / (time.perf_counter() - t0)
        rows.append(dict(ef_search=ef, gpu_qps=qps, recall_at_10=recall_vs(gt_ids, ids), cpu_qps=cpu_qps, cpu_sample=ns,
                         recall_at_10_cpu=recall_vs(gt_ids[:ns], cpu_ids), ids_identical_on_sample=bool(np.array_equal(cpu_ids, ids[:ns])),
                         ratio=qps / cpu_qps))

    def qps_at(target, key):
        # log-linear interpolation of QPS over recall between the two sweep points that bracket the target
        pts = sorted(rows, key=lambda r: r["recall_at_10"])
        for lo, hi in zip(pts, pts[1:]):
            if lo["recall_at_10"] <= target <= hi["recall_at_10"] and hi["recall_at_10"] > lo["recall_at_10"]:
                w = (target - lo["recall_at_10"]) / (hi["recall_at_10"] - lo["recall_at_10"])
                return float(np.exp((1 - w) * np.log(lo[key]) + w * np.log(hi[key])))
        return None

    out = dict(dataset=label, points=rows)
    for target in (0.9, 0.95):
        gq, cq = qps_at(target, "gpu_qps"), qps_at(target, "cpu_qps")
        out["qps_at_recall_%.2f" % target] = dict(gpu=gq, cpu=cq, ratio=(gq / cq) if gq and cq else None) if gq else None
    return out


def broadcast_graph(g, rank, dev):
    """Rank 0's graph arrays to every rank over NCCL (the build does not shard: one rank builds, the replicas receive)."""
    import torch
    import torch.distributed as dist
    meta = [None]
    if rank == 0:
        meta[0] = {k: (g[k].shape, str(g[k].dtype)) for k in ("rows", "level", "adj0", "upper_base", "adj_up")}
        meta[0].update(entry=int(g["entry"]), n=int(g["n"]), source=g["source"])
    dist.broadcast_object_list(meta, src=0)
    m = meta[0]
    out = dict(entry=m["entry"], n=m["n"], source=m["source"] + (" (broadcast from rank 0)" if rank else ""), build_vec_per_s_cpu=None)
    for k in ("rows", "level", "adj0", "upper_base", "adj_up"):
        shape, dt = m[k]
        nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
        if rank == 0:
            t = torch.from_numpy(np.ascontiguousarray(g[k]).view(np.uint8).reshape(-1)).to(dev)
        else:
            t = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
        if nbytes:
            dist.broadcast(t, src=0)
        out[k] = g[k] if rank == 0 else t[:nbytes].cpu().numpy().view(dt).reshape(shape)
        del t
    torch.cuda.empty_cache()
    return out


def strong_scan_section(a, runner, ix, rank, world, dev, batch, reps=10):
    """One caller batch split over the N ranks and returned in ONE host buffer on rank 0 (what amgettuple-side batching
    over several GPUs would do): every rank scans its slice of the same batch, the slices are all-gathered over NCCL and
    rank 0 copies the result to page-locked host memory.  Wall time around `reps` repetitions, barrier on both sides."""
    import torch
    import torch.distributed as dist
    from pgvector_rx_b200 import sharded as S
    out = {}
    for mult in (1, 8):
        nq = len(batch) * mult
        q_all = np.concatenate([batch] * mult) if mult > 1 else batch
        lo, hi = S.query_shard(nq, rank, world)
        q_u8 = q_all[lo:hi] if q_all.dtype == np.float32 else q_all[lo:hi].view(np.uint8)
        d_q = torch.from_numpy(np.ascontiguousarray(q_u8)).to(dev)
        k = a.k
        ids = torch.empty((hi - lo, k), dtype=torch.int32, device=dev)
        ds = torch.empty((hi - lo, k), dtype=torch.float32, device=dev)
        cnt = torch.empty((hi - lo,), dtype=torch.int32, device=dev)
        host_ids = torch.empty((nq, k), dtype=torch.int32).pin_memory()
        host_ds = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        stream = torch.cuda.current_stream().cuda_stream

        def once():
            ix.search_device(d_q.data_ptr(), hi - lo, k, a.ef, ids.data_ptr(), ds.data_ptr(), cnt.data_ptr(), stream=stream)
            gi, gd = S.gather_query_shards(ids, ds, nq)
            if rank == 0:
                host_ids.copy_(gi, non_blocking=True)
                host_ds.copy_(gd, non_blocking=True)
        for _ in range(3):
            once()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["batch_%d" % nq] = dict(value=nq * reps / float(t.item()), unit="queries/s", ms_per_batch=float(t.item()) / reps * 1e3,
                                    queries_per_rank=hi - lo)
    out["note"] = "strong scaling: one caller batch (10k, and 80k) split over %d GPUs, results gathered (NCCL all-gather) into one " \
                  "page-locked host buffer on rank 0; wall clock, max over ranks" % world
    return out


def run_b200(a):
    import torch
    import torch.distributed as dist
    import pgvector_rx_b200 as P

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (no CPU path exists)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c = a.cfg
    elem, metric, norm = P.OPCLASSES[c["opclass"]]
    prep = (lambda q: P.normalize_rows(elem, a.dim, q)[1]) if norm else None

    # ---- the graph: oracle-built cache (C2) or a GPU build; at N > 1 rank 0 builds and the replicas receive it ----
    build_info = None
    ix = None
    if world == 1 or load_cached_graph(a) is not None:
        g, ix, build_info = load_or_build_graph(a, local, want_index=True)
    else:
        g = None
        if rank == 0:
            g, ix, build_info = load_or_build_graph(a, local, want_index=True)
        g = broadcast_graph(g, rank, dev)
    if ix is None:
        ix = P.HnswIndex(c["opclass"], a.dim, m=a.m, ef_construction=a.efc, capacity=a.n, device=local)
        ix.load_graph_stored(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    nb = 4
    batches = query_batches(a, nb, rank, prep)
    runner = ScanRunner(a, ix, batches, dev, world)
    tstream = runner.tstreams[0]
    torch.cuda.set_stream(tstream)

    # ---- device-resident throughput (`value`): the clock sampler is already streaming when the timed region begins ----
    clocks = ClockSampler(local)
    clocks.start()
    for i in range(a.warmup):
        runner.step(i, a.ef, a.k)
    runner.barrier()
    clocks.mark()
    ms_max = runner.device_timed(a.ef, a.k, a.steps, 0)
    value = world * a.nq * a.steps / (ms_max / 1e3)

    # ---- end to end: inputs start in page-locked HOST memory and results end in page-locked HOST memory, every step.
    # (a) the synchronous host-buffer call pgvrx_search_batch (one batch at a time: copy-free for page-locked buffers,
    #     returns when the results are in the caller's arrays);
    # (b) the same buffers through pgvrx_search_batch_device on alternating streams, two batches in flight -- the
    #     kernels read the queries from and write the results to host memory directly -- which is the e2e figure.
    e2e_sync = world * a.nq * a.steps / runner.wall_timed(lambda i: runner.sync_step(i, a.ef, a.k), a.steps, a.warmup)
    e2e = world * a.nq * a.steps / runner.wall_timed(lambda i: runner.step(i, a.ef, a.k, pinned=True), a.steps, a.warmup)
    clk = clocks.stop()
    # the step that ran last wrote pin_outs[(steps - 1) % n_fly]; every query must have completed (count >= 0)
    last_o = runner.pin_outs[a.k][(a.steps - 1) % runner.n_fly]
    assert int(last_o[2].min()) >= 0, "a query exhausted its scratch in the asynchronous path"
    last_ids = last_o[0].numpy().copy()
    last_batch = batches[(a.steps - 1) % nb]

    # ---- counters -> algorithmic bytes ----
    st = P.SearchStats()
    o0 = runner._bufs(a.k, False)[0]
    ix.search_device(runner.d_q[0].data_ptr(), a.nq, a.k, a.ef, o0[0].data_ptr(), o0[1].data_ptr(), o0[2].data_ptr(),
                     stream=tstream.cuda_stream, stats=st)
    row_bytes = ix.L.pgvrx_row_bytes(elem, a.dim)
    step_s = ms_max / 1e3 / a.steps
    kernel = "hnsw_scan_kernel<%s,%s>" % (["F32", "F16", "BIT"][elem], ["L2SQ", "NEG_IP", "L1", "HAMMING", "JACCARD"][metric])
    roof = scan_roofline(a, ix, st, step_s, kernel)
    if rank == 0 and not a.no_gather_probe:
        roof["gather_ceiling"] = gather_probe_section(ix)
        gw = roof["gather_ceiling"].get("graph_walk")
        if gw:
            roof["frac_of_gather_ceiling"] = roof["achieved"] / max(gw["gb_per_s"], gw["gb_per_s_two_half_phases"])

    sweep = staging_sweep(a, runner) if (a.sweep and rank == 0 and world == 1) else None

    # ---- exact kNN: ground truth for recall, and its own line ----
    if a.no_flat:
        flat_info, gt_ids = None, None
    else:
        fb = last_batch if world == 1 else query_batches(a, 1, 0, prep, seed0=c["query_seed"] + 17 * ((a.steps - 1) % nb))[0]
        flat_info, gt_ids = flat_section(a, ix, g, fb, rank, world, local, dev, tstream)

    extras = {}
    if world > 1:
        b0 = query_batches(a, 1, 0, prep)[0]
        extras["strong_scaling_scan"] = strong_scan_section(a, runner, ix, rank, world, dev, b0)

    if rank == 0:
        import oracle_ffi as O
        runner.collective = False          # from here on rank 0 is alone: its timed loops must not wait for the other ranks
        o = oracle_for(a, g)
        # parity: ids, distance bits and work counters of a sample against the oracle's scan of the same graph
        parity = parity_sample(a, ix, o, last_batch, a.ef, a.k, a.parity_sample)
        base, cpu_ids = cpu_baseline(a, o, last_batch, a.ef, a.cpu_sample)
        nchk = len(cpu_ids)
        parity_ok = bool(np.array_equal(cpu_ids, last_ids[:nchk]))
        if gt_ids is None:
            rec = recall_at_k(a, g, last_batch, last_ids, a.recall_sample)
            rec_cpu = recall_at_k(a, g, last_batch, cpu_ids, min(a.recall_sample, nchk))
        else:
            s_chk = min(64 if a.n * a.dim <= 200_000_000 else 16, a.nq)
            o_gt, _ = O.flat_topk(elem, metric, a.dim, g["rows"], last_batch[:s_chk], a.k, nthreads=host_cores(a))
            flat_info["ids_match_cpu_oracle"] = bool(np.array_equal(o_gt, gt_ids[:s_chk]))
            flat_info["cpu_oracle_sample"] = s_chk
            rec = recall_vs(gt_ids, last_ids)
            rec_cpu = recall_vs(gt_ids[:nchk], cpu_ids)

        if a.config == "c2":
            if not a.no_ef100 and a.ef != 100:
                ms100 = runner.device_timed(100, a.k, max(10, a.steps // 4), 3)
                n100 = max(10, a.steps // 4)
                e100 = a.nq * n100 / runner.wall_timed(lambda i: runner.step(i, 100, a.k, pinned=True), n100, 4)
                st100 = P.SearchStats()
                ix.search_device(runner.d_q[0].data_ptr(), a.nq, a.k, 100, o0[0].data_ptr(), o0[1].data_ptr(), o0[2].data_ptr(),
                                 stream=tstream.cuda_stream, stats=st100)
                ids100, _, _ = ix.search(last_batch, k=a.k, ef_search=100, prepared=True)
                a100 = argparse.Namespace(**vars(a))
                a100.ef = 100
                r100 = scan_roofline(a100, ix, st100, ms100 / 1e3 / n100, kernel)
                cpu100, cpu_ids100 = cpu_baseline(a, o, last_batch, 100, 0, budget_s=6.0)
                extras["ef_search_100"] = dict(value=a.nq * n100 / (ms100 / 1e3), unit="queries/s", ms_per_step=ms100 / n100,
                                               e2e=dict(value=e100, unit="queries/s"), roofline=r100,
                                               recall_at_10=recall_vs(gt_ids, ids100) if gt_ids is not None else None,
                                               parity=parity_sample(a, ix, o, last_batch, 100, a.k, a.parity_sample),
                                               ids_match_cpu_oracle=bool(np.array_equal(cpu_ids100, ids100[:len(cpu_ids100)])),
                                               cpu_baseline=cpu100)
            if not a.no_recall_curve and gt_ids is not None:
                efs = (40, 100, 200, 400, 1000)
                curves = [ef_sweep(a, ix, o, runner, gt_ids, last_batch, efs, "gaussian: iid N(0,1), the C2 dataset (no neighbourhood structure in 128 dimensions)")]
                extras["recall_curve"] = curves

        build = dict(cpu_oracle_vectors_per_s=g.get("build_vec_per_s_cpu"), gpu_vectors_per_s=None)
        if build_info:
            build.update(build_info)
        ix_built = None
        if a.config == "c2" and a.build_rows > 0 and not build_info:
            ix.close()
            ix = None
            n_ins = a.insert_rows if a.insert_rows >= 0 else (20000 if (world == 1 and a.build_rows >= a.n) else 0)
            if a.build_rows < a.n:
                n_ins = 0                      # the oracle's graph g (the comparison base) holds all a.n rows
            b_info, ix_built = gpu_build_sample(a, local, n_ins)
            build.update(b_info)
            if ix_built is not None:
                # the scan oracle is no longer needed for queries: let it take the same inserts
                if build.get("adjacency_identical_to_oracle", g.get("source") == "gpu-build"):
                    build["insert"] = insert_sample(a, ix_built, o, n_ins)
                    if a.vacuum_rows > 0 and build["insert"]["graph_identical_to_oracle"]:
                        build["vacuum"] = vacuum_sample(a, ix_built, o, a.vacuum_rows)
                ix_built.close()
        del o

        if a.config == "c2" and not a.no_recall_curve and world == 1 and "recall_curve" in extras:
            extras["recall_curve"].append(lowdim_curve(a, local, dev))

        if c.get("iterative") and ix is not None:
            o = oracle_for(a, g)
            extras["iterative_scan"] = [iterative_section(a, ix, o, last_batch, f, a.iter_queries if f < 500 else max(100, a.iter_queries // 4))
                                        for f in c["filters"]]
            del o
        if c.get("rerank") and ix is not None:
            extras["rerank"] = rerank_section(a, ix, None, last_batch, gt_ids)

        out = dict(metric=metric_name(a), value=value, unit="queries/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                   ms_per_step=ms_max / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=c["dtype"],
                   data="synthetic", recall_at_10=rec, recall_at_10_cpu=rec_cpu, ids_match_cpu_oracle=parity_ok, parity=parity,
                   config=dict(workload=workload_text(a), n=a.n, dim=a.dim, queries_per_step_per_gpu=a.nq,
                               parallelism="query-sharded x%d, graph replicated" % world, batches_in_flight=2, graph=g["source"],
                               full_size=bool(a.full_size),
                               l2="working set (graph + vectors %.2f GB, %.1f GB gathered per step) exceeds L2; batches rotate"
                                  % ((a.n * (row_bytes + 2 * a.m * 4)) / 1e9, roof["bytes_per_launch"] / 1e9)),
                   clocks=clk,
                   e2e=dict(value=e2e, unit="queries/s", h2d_bytes_per_step=a.nq * row_bytes,
                            d2h_bytes_per_step=a.nq * (a.k * 8 + 4), sync_call_value=e2e_sync,
                            note="pgvrx_search_batch_device on page-locked host buffers (queries read from / results written "
                                 "to host memory by the kernels), two batches in flight; sync_call_value = one blocking "
                                 "pgvrx_search_batch call per step"),
                   gpu_launches=3 * a.steps,
                   roofline=roof,
                   cpu_baseline=base,
                   build=build,
                   flat_knn=flat_info)
        out.update(extras)
        if sweep:
            out["staging_sweep"] = sweep
        print(json.dumps(out), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if ix is not None:
        ix.close()


def lowdim_curve(a, local, dev):
    """The same sweep on 1M x 128 rows of low intrinsic dimension (16 latent dimensions + noise), where the graph reaches
    the recall an ANN index is deployed at: GPU-built graph (the oracle imports it), both arms on the same graph."""
    import pgvector_rx_b200 as P
    import torch
    from pgvector_rx_b200 import datasets as D
    rows = D.low_dim_f32(a.n, a.dim, 102)
    lv = D.hnsw_levels(a.n, a.m, a.cfg["level_seed"])
    ix = P.HnswIndex(a.cfg["opclass"], a.dim, m=a.m, ef_construction=a.efc, capacity=a.n, device=local)
    t0 = time.perf_counter()
    st = ix.build(rows, lv)
    dt = time.perf_counter() - t0
    g = ix.export_graph()
    g = dict(rows=g["rows"], level=g["level"], adj0=g["adj0"], upper_base=g["upper_base"], adj_up=g["adj_up"], entry=g["entry"], n=g["n"])
    o = oracle_for(a, g)
    batches = [D.low_dim_f32(a.nq, a.dim, 103 + 17 * i) for i in range(2)]
    runner = ScanRunner(a, ix, batches, dev, 1)
    gt_ids, _ = ix.flat_topk(batches[0], k=a.k)
    out = ef_sweep(a, ix, o, runner, gt_ids, batches[0], (10, 20, 40, 100, 200, 400),
                   "lowdim: 16 latent N(0,1) dimensions through a fixed random 16x128 basis + 0.05 N(0,1) noise per coordinate")
    out["build"] = dict(gpu_vectors_per_s=a.n / dt, rounds=st["n_rounds"], graph="gpu-build (pgvrx_index_build)")
    ix.close()
    return out


def main():
    global _JSON_OUT
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner, ...) goes to stderr
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    # the CPU legs time the oracle built -O3 -march=native on this box; results are identical to
    # the portable build the tests use, and every comparison below re-checks that against the GPU
    os.environ.setdefault("PGVRX_ORACLE_NATIVE", "1")
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
