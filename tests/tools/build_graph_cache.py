"""Build a reference-exact HNSW graph with the CPU oracle and cache it on disk.

The oracle's build is the reference's sequential ambuild loop (~500 vectors/s/core), so a
1M-row graph takes the better part of an hour.  The cached graph lets tests and bench.py compare
the GPU search against the oracle on the SAME graph without paying that build every time.
Output: <out>/{level.npy, adj0.npy, upper_base.npy, adj_up.npy, meta.json}; vectors are
regenerated from the seed, never stored.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_ffi as O  # noqa: E402
from pgvector_rx_b200 import datasets as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--efc", type=int, default=64)
    ap.add_argument("--data-seed", type=int, default=2)
    ap.add_argument("--level-seed", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "data_cache", "c2_1m"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    X = D.gaussian_f32(a.n, a.dim, a.data_seed)
    lv = D.hnsw_levels(a.n, a.m, a.level_seed)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, a.dim, m=a.m, ef_construction=a.efc, tie_mode=O.TIE_TOTAL_ORDER)
    t0 = time.time()
    step = 20000
    for s in range(0, a.n, step):
        o.build(X[s:s + step], lv[s:s + step], base_tid=s)
        el = time.time() - t0
        print(f"{s + step}/{a.n} {el:.0f}s {(s + step) / el:.0f} vec/s", flush=True)
    g = o.export_graph()
    for k in ("level", "adj0", "upper_base", "adj_up"):
        np.save(os.path.join(a.out, k + ".npy"), g[k])
    meta = dict(n=int(g["n"]), dim=a.dim, m=a.m, efc=a.efc, data_seed=a.data_seed, level_seed=a.level_seed,
                entry=int(g["entry"]), build_seconds=time.time() - t0, counters=o.build_counters(),
                elem="f32", metric="l2sq", tie_mode="total_order")
    json.dump(meta, open(os.path.join(a.out, "meta.json"), "w"), indent=1)
    print("done", meta)


if __name__ == "__main__":
    main()
