"""Kernel time of the iterative scan rounds (K3) on the cached C2 graph: python tools/iter_probe.py [nq] [rounds]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pgvector_rx_b200 as P
from pgvector_rx_b200 import datasets as D

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
C = os.path.join(ROOT, "data_cache", "c2_1m")
meta = json.load(open(os.path.join(C, "meta.json")))
g = {k: np.load(os.path.join(C, k + ".npy")) for k in ("level", "adj0", "upper_base", "adj_up")}
rows = D.gaussian_f32(meta["n"], meta["dim"], meta["data_seed"])
ix = P.HnswIndex("vector_l2_ops", meta["dim"], m=meta["m"], ef_construction=meta["efc"], capacity=meta["n"])
ix.load_graph(rows, g["level"], g["adj0"], g["upper_base"], g["adj_up"], meta["entry"])
q = D.gaussian_f32(nq, meta["dim"], 3)
for mode, name in ((P.ITER_RELAXED, "relaxed"), (P.ITER_STRICT, "strict")):
    sc = ix.scan_begin(q, ef_search=40, iterative_scan=mode)
    prev = 0.0
    for r in range(rounds):
        ids, dists, counts, done = sc.next()
        ms = sc.totals["kernel_ms"] - prev
        prev = sc.totals["kernel_ms"]
        print("%s round %d: kernel %.2f ms (%.2f M queries/s), emitted %.1f per query, n_dist so far %d"
              % (name, r, ms, nq / ms / 1e3, counts.mean(), sc.totals["n_dist"]), flush=True)
    sc.close()
ids, dists, counts, st = ix.search(q, k=40, ef_search=40, with_stats=True)
print("one-shot scan (K1) for comparison: kernel %.2f ms" % st["kernel_ms"])
