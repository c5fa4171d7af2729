"""Build n rows of the C2 dataset with pgvrx_index_build, report vectors/s and round statistics, and
(optionally) compare the adjacency with the cached oracle graph of the same size."""
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
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0, help="report progress every CHUNK rows (appending calls)")
    ap.add_argument("--compare", default="", help="data_cache dir of an oracle graph built from the same rows")
    ap.add_argument("--config", default="c2", help="BASELINE config whose rows / options are used (bench.py CONFIGS)")
    a = ap.parse_args()
    import bench
    c = bench.CONFIGS[a.config]
    ba = argparse.Namespace(cfg=c, dim=c["dim"])
    X = bench.raw_rows(ba, a.n, c["data_seed"])
    lv = D.hnsw_levels(a.n, c["m"], c["level_seed"])
    ix = P.HnswIndex(c["opclass"], c["dim"], m=c["m"], ef_construction=c["efc"], capacity=a.n)
    chunk = a.chunk or a.n
    t0 = time.time()
    tot = {}
    for s in range(0, a.n, chunk):
        st = ix.build(X[s:s + chunk], lv[s:s + chunk], heap_tids=np.arange(s, min(a.n, s + chunk)), host_threads=a.threads)
        for k, v in st.items():
            tot[k] = tot.get(k, 0) + v if k != "host_threads" else v
        el = time.time() - t0
        print("%8d rows  %.1fs  %.0f vec/s overall | this chunk: %.0f vec/s, %.2f commits/round, %.2f searches/commit, gpu %.0f%% of time"
              % (s + chunk, el, (s + chunk) / el, chunk / (st["total_ms"] / 1e3), chunk / max(st["n_rounds"], 1),
                 st["n_searched"] / chunk, 100 * st["gpu_ms"] / st["total_ms"]), flush=True)
    print(json.dumps(tot))
    if a.compare:
        g = ix.export_graph()
        meta = json.load(open(os.path.join(a.compare, "meta.json")))
        assert meta["n"] == a.n
        ok = True
        for k in ("level", "adj0", "upper_base", "adj_up"):
            w = np.load(os.path.join(a.compare, k + ".npy"))
            same = np.array_equal(g[k], w)
            ok &= same
            print("compare %-10s %s" % (k, "identical" if same else "DIFFERENT"))
        print("entry", g["entry"], meta["entry"], "=> adjacency %s the oracle's sequential build" % ("IDENTICAL to" if ok and g["entry"] == meta["entry"] else "DIFFERS from"))


if __name__ == "__main__":
    main()
