"""Insert rows into the cached 1M-row C2 graph with pgvrx_index_insert (aminsert's rules) and report where the time goes.
Run with PGVRX_BUILD_DEBUG=1 to get the per-phase timers of the insert loop on stderr."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import pgvector_rx_b200 as P  # noqa: E402
from pgvector_rx_b200 import datasets as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cache", default=os.path.join(ROOT, "data_cache", "c2_1m"))
    ap.add_argument("--rows", type=int, default=20000)
    ap.add_argument("--calls", type=int, default=2)
    a = ap.parse_args()
    meta = json.load(open(os.path.join(a.cache, "meta.json")))
    n = meta["n"]
    X = D.gaussian_f32(n, meta["dim"], meta["data_seed"])
    g = {k: np.load(os.path.join(a.cache, k + ".npy")) for k in ("level", "adj0", "upper_base", "adj_up")}
    ix = P.HnswIndex("vector_l2_ops", meta["dim"], m=meta["m"], ef_construction=meta["efc"], capacity=n + a.rows)
    ix.load_graph(X, g["level"], g["adj0"], g["upper_base"], g["adj_up"], meta["entry"])
    R = D.gaussian_f32(a.rows, meta["dim"], 12)
    lv = D.hnsw_levels(a.rows, meta["m"], 13)
    step = (a.rows + a.calls - 1) // a.calls
    for s in range(0, a.rows, step):
        t0 = time.perf_counter()
        st = ix.insert(R[s:s + step], lv[s:s + step], heap_tids=np.arange(n + s, n + min(a.rows, s + step), dtype=np.uint64))
        dt = time.perf_counter() - t0
        c = min(a.rows, s + step) - s
        print("call %d: %d rows in %.3f s = %.0f vec/s | rounds %d (%.1f commits/round, %.2f ms/round), searches %d, conflicts %d, "
              "big retries %d, gpu_ms %.0f kernel_ms %.0f total_ms %.0f"
              % (s // step, c, dt, c / dt, st["n_rounds"], c / max(st["n_rounds"], 1), st["total_ms"] / max(st["n_rounds"], 1),
                 st["n_searched"], st["n_conflicts"], st["n_big_retries"], st["gpu_ms"], st["kernel_ms"], st["total_ms"]), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
