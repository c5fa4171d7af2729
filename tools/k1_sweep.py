"""K1 tuning sweep on the cached C2 graph (one GPU): device-resident QPS (two batches in flight, CUDA events) for a set of
environment knobs of the scan launch, and the gather-ceiling probe over its geometry.  Prints one JSON object.

    python tools/k1_sweep.py [--ef 40] [--steps 60]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ef", type=int, default=40)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    x = ap.parse_args()
    sys.argv = ["bench.py", "--config", x.config, "--ef", str(x.ef)] + (["--n", str(x.n)] if x.n else [])
    a = bench.parse()
    import torch
    import pgvector_rx_b200 as P
    dev = torch.device("cuda", 0)
    g, ix, _ = bench.load_or_build_graph(a, 0, want_index=True)
    if ix is None:
        ix = P.HnswIndex(a.cfg["opclass"], a.dim, m=a.m, ef_construction=a.efc, capacity=a.n)
        ix.load_graph(g["rows"], g["level"], g["adj0"], g["upper_base"], g["adj_up"], g["entry"])
    elem, metric, norm = P.OPCLASSES[a.cfg["opclass"]]
    prep = (lambda q: P.normalize_rows(elem, a.dim, q)[1]) if norm else None
    runner = bench.ScanRunner(a, ix, bench.query_batches(a, 4, 0, prep), dev, 1)
    out = dict(scan=[], gather=[])
    knobs = [dict()] + [dict(PGVRX_STAGE_ROWS=str(r)) for r in (6, 8, 10, 11, 12, 14, 16, 17, 18, 20, 22, 24, 32)] + \
            [dict(PGVRX_CHUNK_BYTES=str(cb)) for cb in (256, 128, 384)] + \
            [dict(PGVRX_STAGE_ROWS="16", PGVRX_NO_SPEC_GATHER="1"), dict(PGVRX_STAGE_ROWS="11", PGVRX_NO_SPEC_GATHER="1"),
             dict(PGVRX_STAGE_ROWS="16", PGVRX_SCAN_WARPS="2"), dict(PGVRX_STAGE_ROWS="11", PGVRX_SCAN_WARPS="2"), dict()]
    if x.quick:
        knobs = [dict(), dict(PGVRX_SPEC_GATHER="2"), dict(PGVRX_SPEC_GATHER="0"), dict(PGVRX_SPEC_GATHER="1"), dict(PGVRX_SPEC_GATHER="2"), dict(),
                 dict(PGVRX_CHUNK_BYTES="256"), dict(PGVRX_CHUNK_BYTES="256", PGVRX_SPEC_GATHER="0")]
    for kn in knobs:
        for k_, v in kn.items():
            os.environ[k_] = v
        ms = runner.device_timed(a.ef, a.k, x.steps, 6)
        out["scan"].append(dict(knobs=kn, qps=a.nq * x.steps / (ms / 1e3), ms_per_step=ms / x.steps))
        for k_ in kn:
            del os.environ[k_]
    for mode in (() if x.no_gather else (1, 0)):
        for wps, rows, depth in ((0, 32, 1), (6, 32, 1), (3, 32, 1), (0, 16, 2), (0, 16, 1), (0, 32, 2), (0, 8, 4)):
            for consume in (1, 0):
                try:
                    r = ix.gather_probe(mode=mode, warps_per_sm=wps, rows_per_step=rows, depth=depth, steps=192, consume=consume)
                    out["gather"].append(dict(mode="graph_walk" if mode else "uniform", rows_per_phase=rows, phases_in_flight=depth,
                                              consume=consume, **r))
                except Exception as exc:
                    out["gather"].append(dict(mode=mode, rows=rows, depth=depth, error=repr(exc)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
