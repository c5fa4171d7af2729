#!/usr/bin/env python
"""bench.py -- QPS @ recall@10 of the batched HNSW index scan, 1M x 128 fp32 vector_l2_ops, m=16,
ef_construction=64, ef_search=40, 10k-query batches (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step = one pass of the hot path (get_scan_items for every query of one 10k batch) on each GPU.
Queries shard across GPUs against a replicated graph (weak scaling: 10k queries per GPU per step,
no data-path collective).  Prints ONE JSON line on rank 0.

The graph is the reference-exact graph built by the CPU oracle's sequential ambuild loop; because
that takes ~45 minutes for 1M rows it is cached under data_cache/ (tests/tools/build_graph_cache.py) and
only rebuilt when the cache is absent.  Vectors and queries are regenerated from seeds.
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
METRIC = "QPS@recall@10 and build vectors/s, 1M×128 L2 HNSW, 1/2/4/8 B200"


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
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--efc", type=int, default=64)
    ap.add_argument("--ef", type=int, default=40)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--cache", default=os.path.join(ROOT, "data_cache"))
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--recall-sample", type=int, default=200)
    ap.add_argument("--build-rows", type=int, default=-1,
                    help="rows of the same dataset built on the GPU for the build vectors/s figure (0 = skip)")
    ap.add_argument("--insert-rows", type=int, default=-1,
                    help="rows inserted into the finished index through aminsert's rules (-1 = 20000 when the whole dataset was built at N=1, else 0)")
    ap.add_argument("--vacuum-rows", type=int, default=0,
                    help="after the insert sample: delete this many rows of the built index and VACUUM it on the GPU and in the "
                         "oracle, comparing the graphs (0 = skip; every deleted row makes ~2m live elements need a repair search)")
    ap.add_argument("--threads", type=int, default=0, help="host threads for the CPU legs (0 = the cores this process may use)")
    ap.add_argument("--no-flat", action="store_true", help="skip the exact flat-kNN section (recall then uses the oracle's brute force on --recall-sample queries)")
    a = ap.parse_args()
    if a.build_rows < 0:
        # the build does not shard (replicas only): the full 1M-row build is measured at N = 1, larger N keep a 100k-row
        # sample of it on rank 0 so that the scaling runs stay short
        a.build_rows = 1_000_000 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 100_000
    return a


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def load_or_build_graph(a):
    """Returns dict(rows, level, adj0, upper_base, adj_up, entry, n) of the oracle-built graph."""
    from pgvector_rx_b200 import datasets as D
    rows = D.gaussian_f32(a.n, a.dim, 2)
    tag = "c2_1m" if (a.n, a.dim, a.m, a.efc) == (1_000_000, 128, 16, 64) else "g_%d_%d_%d_%d" % (a.n, a.dim, a.m, a.efc)
    d = os.path.join(a.cache, tag)
    meta_p = os.path.join(d, "meta.json")
    if os.path.exists(meta_p):
        meta = json.load(open(meta_p))
        if (meta["n"], meta["dim"], meta["m"], meta["efc"], meta["data_seed"], meta["level_seed"]) == (a.n, a.dim, a.m, a.efc, 2, 4):
            g = {k: np.load(os.path.join(d, k + ".npy")) for k in ("level", "adj0", "upper_base", "adj_up")}
            g.update(rows=rows, entry=meta["entry"], n=meta["n"], source="cache:" + tag,
                     build_vec_per_s_cpu=meta["n"] / meta["build_seconds"])
            return g
    lv = D.hnsw_levels(a.n, a.m, 4)
    g = None
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        # the GPU build produces exactly the oracle's graph (asserted at 1M rows whenever the cache exists); without a
        # cache it is the quick way to get the reference-exact graph.  The CPU build rate is then unknown.
        log("graph cache %s missing: building %d rows on the GPU" % (d, a.n))
        import pgvector_rx_b200 as P
        bx = P.HnswIndex("vector_l2_ops", a.dim, m=a.m, ef_construction=a.efc, capacity=a.n,
                         device=int(os.environ.get("LOCAL_RANK", "0")))
        t0 = time.time()
        bx.build(rows, lv)
        dt = time.time() - t0
        g = bx.export_graph()
        bx.close()
        g = dict(rows=rows, level=g["level"], adj0=g["adj0"], upper_base=g["upper_base"], adj_up=g["adj_up"], entry=g["entry"],
                 n=g["n"], source="gpu-build", build_vec_per_s_cpu=None)
        return g
    log("graph cache %s missing: building %d rows with the CPU oracle (slow)" % (d, a.n))
    import oracle_ffi as O
    o = O.Oracle(O.ELEM_F32, O.L2SQ, a.dim, m=a.m, ef_construction=a.efc, tie_mode=O.TIE_TOTAL_ORDER)
    t0 = time.time()
    o.build(rows, lv)
    dt = time.time() - t0
    g = o.export_graph()
    g.update(source="oracle-build", build_vec_per_s_cpu=a.n / dt)
    try:
        os.makedirs(d, exist_ok=True)
        for k in ("level", "adj0", "upper_base", "adj_up"):
            np.save(os.path.join(d, k + ".npy"), g[k])
        json.dump(dict(n=int(g["n"]), dim=a.dim, m=a.m, efc=a.efc, data_seed=2, level_seed=4, entry=int(g["entry"]),
                       build_seconds=dt), open(meta_p, "w"))
    except OSError:
        pass
    return g


def query_batches(a, count, rank):
    from pgvector_rx_b200 import datasets as D
    return [D.gaussian_f32(a.nq, a.dim, 3 + 1000 * rank + 17 * i) for i in range(count)]


def oracle_for(a, g):
    import oracle_ffi as O
    o = O.Oracle(O.ELEM_F32, O.L2SQ, a.dim, m=a.m, ef_construction=a.efc, tie_mode=O.TIE_TOTAL_ORDER)
    o.import_graph(g)
    return o


def cpu_baseline(a, g, queries, sample=0, budget_s=12.0):
    """The oracle (CPU restatement of the reference's scan) on all host cores: passes over the same
    10k-query batch (or `sample` queries of it) until about budget_s seconds of wall time are used."""
    o = oracle_for(a, g)
    cores = host_cores(a)
    q = queries[:sample] if sample else queries
    o.search_batch(q[: 8 * cores], ef_search=a.ef, k=a.k, nthreads=cores)      # warm caches / threads
    done, t0, ids = 0, time.perf_counter(), None
    while True:
        r = o.search_batch(q, ef_search=a.ef, k=a.k, nthreads=cores)
        ids = r[0] if ids is None else ids
        done += len(q)
        dt = time.perf_counter() - t0
        if dt >= budget_s or sample:
            break
    return dict(value=done / dt, unit="queries/s", cores=cores, kind="port",
                sample="%d queries (%d passes over %d of the batch), ef_search=%d, k=%d, %d threads, %.1f s"
                       % (done, done // len(q), len(q), a.ef, a.k, cores, dt)), ids, o


def gpu_build_sample(a, local, n_insert=0):
    """Index build (CREATE INDEX) of the first --build-rows rows of the C2 dataset through
    pgvrx_index_build: GPU-batched search_layer + host heuristic.  When an oracle graph of exactly that
    prefix is cached, the adjacency is compared with it (it must be identical)."""
    import pgvector_rx_b200 as P
    from pgvector_rx_b200 import datasets as D
    n = min(a.build_rows, a.n)
    X = D.gaussian_f32(a.n, a.dim, 2)[:n] if n < a.n else D.gaussian_f32(a.n, a.dim, 2)
    lv = D.hnsw_levels(a.n, a.m, 4)[:n]
    ix = P.HnswIndex("vector_l2_ops", a.dim, m=a.m, ef_construction=a.efc, capacity=n + n_insert, device=local)
    t0 = time.perf_counter()
    st = ix.build(X, lv)
    dt = time.perf_counter() - t0
    out = dict(gpu_vectors_per_s=n / dt, rows=n, seconds=dt, rounds=st["n_rounds"], searches=st["n_searched"],
               conflicts=st["n_conflicts"], kernel_ms=st["kernel_ms"], host_threads=st["host_threads"],
               commits_per_round=n / max(st["n_rounds"], 1), pairwise_on_gpu=st["n_pair_gpu"], host_distance_evals=st["n_host_dist"],
               note="%s rows of the C2 dataset (same seeds and levels as the oracle's build); exact sequential semantics, "
                    "see DESIGN.md" % ("all %d" % n if n == a.n else "first %d" % n))
    tag = "c2_1m" if (n, a.dim, a.m, a.efc) == (1_000_000, 128, 16, 64) else "g_%d_%d_%d_%d" % (n, a.dim, a.m, a.efc)
    d = os.path.join(a.cache, tag)
    if os.path.exists(os.path.join(d, "meta.json")):
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
    s = min(sample, len(queries))
    gt, _ = O.flat_topk(O.ELEM_F32, O.L2SQ, a.dim, g["rows"], queries[:s], a.k, nthreads=host_cores(a))
    hit = sum(len(set(got_ids[i].tolist()) & set(gt[i].tolist())) for i in range(s))
    return hit / float(s * a.k)


def recall_vs(gt_ids, got_ids):
    n, k = got_ids.shape
    hit = sum(len(set(got_ids[i].tolist()) & set(gt_ids[i].tolist())) for i in range(n))
    return hit / float(n * k)


def flat_section(a, ix, g, batch, rank, world, local, dev, tstream):
    """Exact flat kNN: ground truth for recall and its own throughput line.

    N = 1: the whole table on one GPU (3xTF32 tcgen05 candidates + exact re-rank, certified).
    N > 1: the base set is sharded by rows, every rank sees the same query batch, local top-k lists are
    all-gathered over NCCL and merged by the K6 kernel on every rank (strong scaling of one batch)."""
    import torch
    import torch.distributed as dist
    import pgvector_rx_b200 as P
    from pgvector_rx_b200 import sharded as S
    stream = tstream.cuda_stream
    d_q = torch.from_numpy(batch).to(dev)
    nq, k = len(batch), a.k
    st = P.FlatStats()
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    ds = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out = {}
    if world == 1:
        fx = ix
        lo = 0
        run = lambda: fx.flat_topk_device(d_q.data_ptr(), nq, k, ids.data_ptr(), ds.data_ptr(), stream=stream)
        result = lambda: ids
    else:
        lo, hi = S.shard_bounds(a.n, world)[rank]
        fx = P.HnswIndex("vector_l2_ops", a.dim, m=a.m, ef_construction=a.efc, capacity=hi - lo, device=local)
        fx.load_rows(g["rows"][lo:hi])
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
    reps = 10
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
    ms = float(t.item())
    fx.flat_topk_device(d_q.data_ptr(), nq, k, ids.data_ptr(), ds.data_ptr(), id_base=lo, stats=st)
    useful = 2.0 * nq * a.n * a.dim
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf32_est = peaks.get("bf16_tflops", 2250.0 * 0.73) / 2.0
    # the tf32 peak itself is not in MEASURED_PEAKS.json: probe it here with a library GEMM (cuBLAS through torch, used
    # for nothing but this denominator), best of 10 like the driver's bf16 figure
    tf32_meas = None
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
        tf32_meas = 2.0 * 8192 ** 3 / (best / 1e3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = old_flag
        del ma, mb, mc
    except Exception as exc:                                    # the probe is optional
        log("tf32 peak probe failed: %r" % (exc,))
    tf32_peak = tf32_meas if tf32_meas else tf32_est
    issued = useful * (3 if st.path >= 1 else 1) / world
    out = dict(value=nq / (ms / 1e3), unit="queries/s", ms_per_batch=ms, k=k, rows=a.n, queries=nq, n_gpus=world,
               path="3xTF32 tcgen05 candidates + exact re-rank (certified)" if st.path >= 1 else "sequential f32 chains (CUDA cores)",
               candidates_per_query=int(st.n_candidates // max(nq, 1)), uncertified=int(st.n_uncertified),
               parallelism="single table" if world == 1 else "base rows sharded x%d, NCCL all-gather of %d B/rank + K6 merge" % (world, nq * k * 8),
               roofline=dict(bound="tensor", achieved=issued / (ms / 1e3) / 1e12, peak=tf32_peak, unit="TFLOP/s",
                             frac=issued / (ms / 1e3) / 1e12 / tf32_peak, traffic=None,
                             peak_source="torch.matmul TF32 8192^3, best of 10, this run" if tf32_meas else "measured dense bf16 / 2",
                             peak_bf16_half=tf32_est, frac_of_nominal_1125=issued / (ms / 1e3) / 1e12 / 1125.0,
                             note="issued tf32 flops per GPU (3 products x 2*Q*N*D / n_gpus) / CUDA-event time of the whole call "
                                  "(operand split, GEMM + top-k, re-rank, certification%s)"
                                  % (", all-gather, merge" if world > 1 else "")))
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

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the scan kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get("dram_bytes_per_launch")
    return None


# ---------------------------------------------------------------------------------------------
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = load_or_build_graph(a)
    batches = query_batches(a, 1, 0)
    o = oracle_for(a, g)
    cores = host_cores(a)
    per_step = a.nq if cores >= 8 else min(a.nq, 400 * cores)   # the whole 10k batch per step unless the box is tiny
    q = batches[0][:per_step]
    for _ in range(max(a.warmup, 1)):
        o.search_batch(q[: 4 * cores], ef_search=a.ef, k=a.k, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ids, _, _ = o.search_batch(q, ef_search=a.ef, k=a.k, nthreads=cores)
    dt = time.perf_counter() - t0
    qps = per_step * a.steps / dt
    rec = recall_at_k(a, g, q, ids, a.recall_sample)
    sample = "%d queries/step of the 10k batch, %d threads" % (per_step, cores)
    out = dict(impl="reference", metric=METRIC, value=qps, unit="queries/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
               ms_per_step=dt / a.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
               data="synthetic", recall_at_10=rec,
               config=dict(workload="C2: %sx%d fp32 vector_l2_ops m=%d ef_construction=%d ef_search=%d k=%d, %d-query batches"
                                    % ("1M" if a.n == 1_000_000 else str(a.n), a.dim, a.m, a.efc, a.ef, a.k, a.nq), n=a.n, dim=a.dim, queries_per_step_per_gpu=per_step,
                           parallelism="host threads x%d" % cores, graph=g["source"],
                           note="CPU restatement of pgvector-rx's scan on flat arrays (the Rust/pgrx reference cannot be built here)"),
               build=dict(cpu_oracle_vectors_per_s=g.get("build_vec_per_s_cpu"),
                          note="sequential ambuild restatement, one core (the reference's build is single-threaded per backend), "
                               "measured when the cached graph was built"),
               cpu_baseline=dict(value=qps, unit="queries/s", cores=cores, kind="port", sample=sample),
               e2e=dict(value=qps, unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), file=_JSON_OUT, flush=True)


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    g = load_or_build_graph(a)
    ix = P.HnswIndex("vector_l2_ops", a.dim, m=a.m, ef_construction=a.efc, capacity=a.n, device=local)
    ix.load_graph(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    nb = 4
    batches = query_batches(a, nb, rank)
    dev = torch.device("cuda", local)
    d_q = [torch.from_numpy(b).to(dev) for b in batches]
    # two batches in flight: consecutive steps go to alternating streams (each with its own result buffers; the library
    # keeps one scratch set per stream), so the ragged tail of one persistent launch -- warps finishing their last query
    # at different times -- is filled by the next batch instead of idling the SMs.  torch events are recorded on the
    # streams the library launches on.
    n_fly = 2
    tstreams = [torch.cuda.Stream(device=dev) for _ in range(n_fly)]
    tstream = tstreams[0]
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    outs = [(torch.empty((a.nq, a.k), dtype=torch.int32, device=dev), torch.empty((a.nq, a.k), dtype=torch.float32, device=dev),
             torch.empty((a.nq,), dtype=torch.int32, device=dev)) for _ in range(n_fly)]
    d_ids, d_dists, d_cnt = outs[0]

    def step_device(i):
        o = outs[i % n_fly]
        ix.search_device(d_q[i % nb].data_ptr(), a.nq, a.k, a.ef, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                         stream=tstreams[i % n_fly].cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ----
    for i in range(a.warmup):
        step_device(i)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for i in range(a.steps):
        step_device(i)
    for t in tstreams[1:]:                       # the timed region ends when every stream has drained
        ev = torch.cuda.Event()
        ev.record(t)
        tstream.wait_event(ev)
    e1.record(tstream)
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * a.nq * a.steps / (ms_max / 1e3)

    # ---- end to end: inputs start in page-locked HOST memory and results end in page-locked HOST memory, every step.
    # (a) the synchronous host-buffer call pgvrx_search_batch (one batch at a time: copy-free for page-locked buffers,
    #     returns when the results are in the caller's arrays);
    # (b) the same buffers through pgvrx_search_batch_device on alternating streams, two batches in flight -- the
    #     kernels read the queries from and write the results to host memory directly -- which is the e2e figure.
    pin_q = [torch.from_numpy(b).pin_memory() for b in batches]
    pin_out = [(torch.empty((a.nq, a.k), dtype=torch.int32).pin_memory(), torch.empty((a.nq, a.k), dtype=torch.float32).pin_memory(),
                torch.empty((a.nq,), dtype=torch.int32).pin_memory()) for _ in range(n_fly)]
    pin_ids, pin_d, pin_c = pin_out[0]
    import ctypes as C

    def step_host_sync(i):
        rc = ix.L.pgvrx_search_batch(ix.h, C.c_void_p(pin_q[i % nb].data_ptr()), a.nq, a.ef, a.k, C.c_void_p(pin_ids.data_ptr()),
                                     C.c_void_p(pin_d.data_ptr()), C.c_void_p(pin_c.data_ptr()), None)
        ix._check(rc)

    def step_host_async(i):
        o = pin_out[i % n_fly]
        ix.search_device(pin_q[i % nb].data_ptr(), a.nq, a.k, a.ef, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                         stream=tstreams[i % n_fly].cuda_stream)

    def timed(step_fn):
        for i in range(max(3, a.warmup // 2)):
            step_fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            step_fn(i)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * a.nq * a.steps / float(tt.item())

    e2e_sync = timed(step_host_sync)
    e2e = timed(step_host_async)
    # the step that ran last wrote pin_out[(steps - 1) % n_fly]; every query must have completed (count >= 0)
    last_o = pin_out[(a.steps - 1) % n_fly]
    assert int(last_o[2].min()) >= 0, "a query exhausted its scratch in the asynchronous path"
    last_ids = last_o[0].numpy().copy()
    last_batch = batches[(a.steps - 1) % nb]

    # ---- counters -> algorithmic bytes; parity + recall + CPU baseline on rank 0 ----
    st = P.SearchStats()
    ix.search_device(d_q[0].data_ptr(), a.nq, a.k, a.ef, d_ids.data_ptr(), d_dists.data_ptr(), d_cnt.data_ptr(), stream=stream, stats=st)
    row_bytes = a.dim * 4
    bytes_step = st.n_dist * row_bytes + st.n_expand0 * 2 * a.m * 4 + st.n_expand_upper * a.m * 4
    peak, peak_src = measured_peak()
    step_s = ms_max / 1e3 / a.steps
    achieved = bytes_step / step_s / 1e9
    flat_info, gt_ids = (None, None) if a.no_flat else flat_section(a, ix, g, last_batch if world == 1 else batches_rank0_last(a),
                                                                    rank, world, local, dev, tstream)
    out = None
    if rank == 0:
        build_info = dict(cpu_oracle_vectors_per_s=g.get("build_vec_per_s_cpu"), gpu_vectors_per_s=None)
        ix_built = None
        if a.build_rows > 0:
            ix.close()
            n_ins = a.insert_rows if a.insert_rows >= 0 else (20000 if (world == 1 and a.build_rows >= a.n) else 0)
            if a.build_rows < a.n:
                n_ins = 0                      # the oracle's graph g (the comparison base) holds all a.n rows
            b_info, ix_built = gpu_build_sample(a, local, n_ins)
            build_info.update(b_info)
        base, cpu_ids, o = cpu_baseline(a, g, last_batch, a.cpu_sample)
        if ix_built is not None:
            # the scan oracle is no longer needed for queries: let it take the same inserts
            if build_info.get("adjacency_identical_to_oracle", g.get("source") == "gpu-build"):
                build_info["insert"] = insert_sample(a, ix_built, o, n_ins)
                if a.vacuum_rows > 0 and build_info["insert"]["graph_identical_to_oracle"]:
                    build_info["vacuum"] = vacuum_sample(a, ix_built, o, a.vacuum_rows)
            ix_built.close()
        nchk = len(cpu_ids)
        parity_ok = bool(np.array_equal(cpu_ids, last_ids[:nchk]))
        # recall of the whole batch against the GPU's exact kNN; the exact result itself is cross-checked against the
        # oracle's brute force on a sample (ids must be identical)
        import oracle_ffi as O
        if gt_ids is None:
            rec = recall_at_k(a, g, last_batch, last_ids, a.recall_sample)
            rec_cpu = recall_at_k(a, g, last_batch, cpu_ids, min(a.recall_sample, nchk))
        else:
            s_chk = min(64, a.nq)
            o_gt, _ = O.flat_topk(O.ELEM_F32, O.L2SQ, a.dim, g["rows"], last_batch[:s_chk], a.k, nthreads=host_cores(a))
            flat_info["ids_match_cpu_oracle"] = bool(np.array_equal(o_gt, gt_ids[:s_chk]))
            rec = recall_vs(gt_ids, last_ids)
            rec_cpu = recall_vs(gt_ids[:nchk], cpu_ids)
        out = dict(metric=METRIC, value=value, unit="queries/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                   ms_per_step=ms_max / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                   data="synthetic", recall_at_10=rec, recall_at_10_cpu=rec_cpu, ids_match_cpu_oracle=parity_ok,
                   config=dict(workload="C2: %sx%d fp32 vector_l2_ops m=%d ef_construction=%d ef_search=%d k=%d, %d-query batches"
                                        % ("1M" if a.n == 1_000_000 else str(a.n), a.dim, a.m, a.efc, a.ef, a.k, a.nq), n=a.n, dim=a.dim, queries_per_step_per_gpu=a.nq,
                               parallelism="query-sharded x%d, graph replicated" % world, batches_in_flight=2, graph=g["source"],
                               l2="working set (graph + vectors 0.64 GB, 6.7 GB gathered per step) exceeds L2; batches rotate"),
                   clocks=clk,
                   e2e=dict(value=e2e, unit="queries/s", h2d_bytes_per_step=a.nq * row_bytes,
                            d2h_bytes_per_step=a.nq * (a.k * 8 + 4), sync_call_value=e2e_sync,
                            note="pgvrx_search_batch_device on page-locked host buffers (queries read from / results written "
                                 "to host memory by the kernels), two batches in flight; sync_call_value = one blocking "
                                 "pgvrx_search_batch call per step"),
                   gpu_launches=3 * a.steps,
                   roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, frac_of_nominal_8000=achieved / 8000.0,
                                 traffic=ncu_traffic(),
                                 peak_source=peak_src, kernel="hnsw_scan_kernel<F32,L2SQ>",
                                 bytes_per_launch=int(bytes_step), bytes_per_query=bytes_step / a.nq,
                                 counters=dict(n_dist=int(st.n_dist), n_expand0=int(st.n_expand0),
                                               n_expand_upper=int(st.n_expand_upper), n_retried=int(st.n_retried)),
                                 note="achieved = (n_dist*512 + n_expand0*128 + n_expand_upper*64 bytes per step) / CUDA-event step time, rank 0 counters"),
                   cpu_baseline=base,
                   build=build_info,
                   flat_knn=flat_info)
        print(json.dumps(out), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ix.close()


def batches_rank0_last(a):
    """The batch rank 0 ran last (every rank needs the same queries for the sharded flat path)."""
    from pgvector_rx_b200 import datasets as D
    return D.gaussian_f32(a.nq, a.dim, 3 + 17 * ((a.steps - 1) % 4))


def main():
    global _JSON_OUT
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner, ...) goes to stderr
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
