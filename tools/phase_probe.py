"""Profiling aid: run one 10k-query scan with the phase-timing build and print cycles per phase.

    python -m pgvector_rx_b200.build --timing
    PGVRX_LIB=pgvector_rx_b200/libpgvrx_b200_timing.so python tools/phase_probe.py --n 100000
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import bench  # noqa: E402
import pgvector_rx_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--ef", type=int, default=40)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=3)
    a0 = ap.parse_args()
    a = argparse.Namespace(n=a0.n, dim=128, m=16, efc=64, ef=a0.ef, k=10, nq=a0.nq, cache=os.path.join(ROOT, "data_cache"))
    g = bench.load_or_build_graph(a)
    ix = P.HnswIndex("vector_l2_ops", a.dim, m=a.m, ef_construction=a.efc, capacity=a.n)
    ix.load_graph(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    q = bench.query_batches(a, 1, 0)[0]
    for r in range(a0.reps):
        t0 = time.perf_counter()
        ids, d, c, st = ix.search(q, k=10, ef_search=a.ef, with_stats=True)
        dt = time.perf_counter() - t0
        print("rep %d: kernel %.3f ms (%.2f M qps), host call %.3f ms, n_dist/q %.0f, expansions/q %.1f, retried %d" % (
            r, st["kernel_ms"], a.nq / st["kernel_ms"] / 1e3, dt * 1e3, st["n_dist"] / a.nq,
            (st["n_expand0"] + st["n_expand_upper"]) / a.nq, st["n_retried"]), flush=True)


if __name__ == "__main__":
    main()
