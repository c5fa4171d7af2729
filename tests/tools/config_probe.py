"""BASELINE configs 3-5 at reduced row counts: GPU build, scan QPS, iterative scan, flat kNN, with parity against
the oracle on a sample (the oracle imports the GPU-built graph; adjacency parity of the build itself is covered by
tests/ and by bench.py at 1M rows).

    python tests/tools/config_probe.py c3|c4|c5 [n] [nq]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle_ffi as O  # noqa: E402
import pgvector_rx_b200 as P  # noqa: E402
from pgvector_rx_b200 import datasets as D  # noqa: E402

CFG = {
    "c3": dict(opclass="halfvec_cosine_ops", dim=768, m=32, efc=128, ef=40, n=200_000, seed=5),
    "c4": dict(opclass="vector_ip_ops", dim=96, m=16, efc=64, ef=40, n=1_000_000, seed=7),
    "c5": dict(opclass="bit_hamming_ops", dim=1024, m=16, efc=64, ef=40, n=1_000_000, seed=8),
}


def rows_for(c, n, seed):
    elem = P.OPCLASSES[c["opclass"]][0]
    if elem == P.ELEM_BIT:
        return D.clustered_bits(n, c["dim"], seed)
    x = D.gaussian_f32(n, c["dim"], seed)
    return x.astype(np.float16) if elem == P.ELEM_F16 else x


def main():
    name = sys.argv[1]
    c = CFG[name]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else c["n"]
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
    elem, metric, norm = P.OPCLASSES[c["opclass"]]
    rows, q = rows_for(c, n, c["seed"]), rows_for(c, nq, c["seed"] + 1)
    lv = D.hnsw_levels(n, c["m"], 4)
    ix = P.HnswIndex(c["opclass"], c["dim"], m=c["m"], ef_construction=c["efc"], capacity=n)
    t0 = time.perf_counter()
    st = ix.build(rows, lv)
    t_build = time.perf_counter() - t0
    out = dict(config=name, opclass=c["opclass"], n=len(ix), dim=c["dim"], m=c["m"], efc=c["efc"], ef_search=c["ef"],
               build_vectors_per_s=len(ix) / t_build, build_commits_per_round=len(ix) / max(st["n_rounds"], 1))
    # scan throughput (host-buffer C ABI call, includes H2D / D2H) and parity of a sample against the oracle
    k = 10
    ix.search(q[:256], k=k, ef_search=c["ef"])
    t0 = time.perf_counter()
    ids, dists, counts, sst = ix.search(q, k=k, ef_search=c["ef"], with_stats=True)
    t_scan = time.perf_counter() - t0
    rb = ix.L.pgvrx_row_bytes(elem, c["dim"])
    bytes_q = (sst["n_dist"] * rb + sst["n_expand0"] * 2 * c["m"] * 4 + sst["n_expand_upper"] * c["m"] * 4) / nq
    out.update(scan_qps_e2e=nq / t_scan, scan_kernel_qps=nq / (sst["kernel_ms"] / 1e3), scan_gathered_bytes_per_query=bytes_q,
               scan_gather_GBps=bytes_q * nq / (sst["kernel_ms"] / 1e3) / 1e9, scan_retried=sst["n_retried"])
    g = ix.export_graph()
    o = O.Oracle(elem, metric, c["dim"], m=c["m"], ef_construction=c["efc"], tie_mode=O.TIE_TOTAL_ORDER)
    o.import_graph(g)
    ns = 200
    qn = ix.prepare_queries(q[:ns])
    oi, od, oc = o.search_batch(qn, ef_search=c["ef"], k=k, nthreads=os.cpu_count() or 1)
    out["scan_matches_oracle_on_sample"] = bool(np.array_equal(oi, ids[:ns]) and np.array_equal(od, dists[:ns]))
    # exact kNN and recall
    t0 = time.perf_counter()
    fi, fd, fst = ix.flat_topk(q, k=k, with_stats=True)
    t_flat = time.perf_counter() - t0
    out.update(flat_qps=nq / t_flat, flat_kernel_qps=nq / (fst["kernel_ms"] / 1e3), flat_path=fst["path"],
               recall_at_10=float(np.mean([len(set(ids[i]) & set(fi[i])) / k for i in range(nq)])))
    wi, wd = O.flat_topk(elem, metric, c["dim"], g["rows"], qn[:16], k, nthreads=os.cpu_count() or 1)
    out["flat_matches_oracle_on_sample"] = bool(np.array_equal(wi, fi[:16]) and np.array_equal(wd, fd[:16]))
    if name == "c4":
        # relaxed-order iterative scan with a consumer that keeps id % 50 == 0 and stops after 10 hits
        nqi = 2000
        t0 = time.perf_counter()
        sc = ix.scan_begin(q[:nqi], ef_search=c["ef"], iterative_scan=P.ITER_RELAXED, max_scan_tuples=20000)
        hits = [[] for _ in range(nqi)]
        done = np.zeros(nqi, bool)
        rounds = 0
        while True:
            want = ~done & np.array([len(h) < 10 for h in hits])
            if not want.any():
                break
            ii, dd, cc, dn = sc.next(want.astype(np.uint8))
            rounds += 1
            for i in np.nonzero(want)[0]:
                sel = ii[i, :cc[i]]
                sel = sel[sel % 50 == 0]
                hits[i].extend(sel[: 10 - len(hits[i])].tolist())
            done |= dn
        t_it = time.perf_counter() - t0
        out.update(iterative_filtered_qps=nqi / t_it, iterative_rounds=rounds, iterative_kernel_ms=sc.totals["kernel_ms"])
        sc.close()
        ok = True
        for i in range(8):
            _, eids, _, _ = o.scan_stream(qn[i], ef_search=c["ef"], iterative=O.ITER_RELAXED, max_scan_tuples=20000)
            ok &= [int(e) for e in eids if e % 50 == 0][:10] == hits[i]
        out["iterative_matches_oracle_on_sample"] = bool(ok)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
