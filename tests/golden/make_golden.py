"""Generates tests/golden/*.npz: small seeded cases with the ORACLE's outputs frozen.

    python tests/golden/make_golden.py

The reference itself (Rust / pgrx / PostgreSQL) cannot run in this image, so these are not outputs of the
reference binary: they are outputs of oracle/ (the restatement pinned to the reference's own known-answer
tests by tests/test_oracle_known_answers.py), frozen so that (a) a later change of the oracle that alters
any result is caught on CPU, and (b) the CUDA path is compared on the GPU box against numbers that were
committed before it ran.  Inputs are regenerated from seeds (pgvector_rx_b200.datasets), only graphs'
adjacency and the expected outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import oracle_ffi as O                      # noqa: E402
import pgvector_rx_b200 as P                # noqa: E402
from helpers import make_rows, oracle_index  # noqa: E402

# name -> (opclass, kind, n, dim, nq, m, efc, ef_search, k, seed)
CASES = {
    "c1_vector_l2": ("vector_l2_ops", "f32", 2500, 128, 64, 16, 64, 40, 10, 101),
    "c3_halfvec_cosine": ("halfvec_cosine_ops", "f16", 400, 768, 12, 32, 128, 40, 10, 103),
    "c4_vector_ip": ("vector_ip_ops", "f32", 2000, 96, 32, 16, 64, 40, 10, 104),
    "c5_bit_hamming": ("bit_hamming_ops", "bit", 2000, 1024, 32, 16, 64, 40, 10, 105),
    "vector_l1_small_m": ("vector_l1_ops", "f32", 600, 17, 20, 4, 16, 8, 5, 106),
}


def generate(name):
    opclass, kind, n, dim, nq, m, efc, ef, k, seed = CASES[name]
    elem, metric, norm = P.OPCLASSES[opclass]
    rows = make_rows(kind, n, dim, seed)
    q = make_rows(kind, nq, dim, seed + 1)
    o = oracle_index(opclass, dim, rows, m=m, efc=efc)
    g = o.export_graph()
    qn = np.stack([O.normalize(elem, dim, r)[1] for r in q]) if norm else q
    ids, dists, counts, ctr = o.search_batch(qn, ef_search=ef, k=k, with_counters=True)
    flat_ids, flat_d = O.flat_topk(elem, metric, dim, g["rows"], qn, k)
    out = dict(level=g["level"], adj0=g["adj0"], upper_base=g["upper_base"], adj_up=g["adj_up"],
               entry=np.int64(g["entry"]), n=np.int64(g["n"]), scan_ids=ids, scan_dists=dists, scan_counts=counts,
               counters=np.array([ctr["n_dist"], ctr["n_expand0"], ctr["n_expand_upper"], ctr["n_visited"]], np.int64),
               flat_ids=flat_ids, flat_dists=flat_d)
    # iterative scan: the first 60 elements of the relaxed and strict streams of the first 4 queries
    for mode, tag in ((O.ITER_RELAXED, "relaxed"), (O.ITER_STRICT, "strict")):
        streams = []
        for i in range(4):
            _, eids, ds, _ = o.scan_stream(qn[i], ef_search=ef, iterative=mode, max_scan_tuples=20000, max_emit=60)
            pad = np.full(60, -1, np.int32)
            pad[:len(eids)] = eids
            streams.append(pad)
        out["iter_" + tag] = np.stack(streams)
    return out


# A LIVE index: n0 rows by CREATE INDEX, n1 rows by aminsert (src/index/insert.rs), then the states a vacuum leaves
# behind (deleted / no heap TIDs / stale neighbour tuples, src/index/vacuum.rs) and scans over them.
# name -> (opclass, kind, n0, n1, dim, nq, m, efc, ef_search, k, seed)
LIVE_CASES = {
    "live_vector_l2": ("vector_l2_ops", "f32", 1200, 600, 24, 24, 8, 32, 40, 10, 201),
    "live_bit_hamming": ("bit_hamming_ops", "bit", 900, 500, 256, 16, 6, 24, 20, 10, 202),
}


def live_inputs(name):
    from pgvector_rx_b200 import datasets as D
    opclass, kind, n0, n1, dim, nq, m, efc, ef, k, seed = LIVE_CASES[name]
    rows = make_rows(kind, n0 + n1, dim, seed)
    lv = D.hnsw_levels(n0 + n1, m, seed + 2)
    q = make_rows(kind, nq, dim, seed + 1)
    return rows, lv, q


def live_oracle(name):
    """The oracle after build + insert, then after the seeded vacuum marks; returns (o, graph after insert, marks)."""
    opclass, kind, n0, n1, dim, nq, m, efc, ef, k, seed = LIVE_CASES[name]
    elem, metric, norm = P.OPCLASSES[opclass]
    rows, lv, q = live_inputs(name)
    o = O.Oracle(elem, metric, dim, m=m, ef_construction=efc, tie_mode=O.TIE_TOTAL_ORDER)
    o.build(rows[:n0], lv[:n0])
    o.insert(rows[n0:], lv[n0:], base_tid=n0)
    after_insert = o.export_graph()
    rng = np.random.default_rng(seed)
    perm = rng.permutation(o.n)
    c = o.n // 16
    deleted, no_tid, stale = perm[:c], perm[c:2 * c], perm[2 * c:2 * c + c // 4]
    deleted = deleted[deleted != o.entry]
    o.vacuum_mark(deleted, O.VACUUM_MARK_DELETED)
    o.vacuum_mark(no_tid, O.VACUUM_REMOVE_HEAPTIDS)
    o.vacuum_mark(stale, O.VACUUM_STALE_NEIGHBORS)
    return o, after_insert, q


def generate_live(name):
    opclass, kind, n0, n1, dim, nq, m, efc, ef, k, seed = LIVE_CASES[name]
    elem = P.OPCLASSES[opclass][0]
    o, gi, q = live_oracle(name)
    out = dict(ins_level=gi["level"], ins_adj0=gi["adj0"], ins_upper_base=gi["upper_base"], ins_adj_up=gi["adj_up"],
               ins_entry=np.int64(gi["entry"]), ins_n=np.int64(gi["n"]))
    # what the scan sees after the vacuum: the relation's pages through the product's exporter
    gv = P.export_pages(o.write_pages(), elem)
    assert np.array_equal(gv["flags"], o.element_flags())
    out.update(vac_adj0=gv["adj0"], vac_adj_up=gv["adj_up"], vac_flags=gv["flags"])
    ids, dists, counts, ctr = o.search_batch(q, ef_search=ef, k=k, with_counters=True)
    out.update(scan_ids=ids, scan_dists=dists, scan_counts=counts,
               counters=np.array([ctr["n_dist"], ctr["n_expand0"], ctr["n_expand_upper"], ctr["n_visited"]], np.int64))
    for mode, tag in ((O.ITER_RELAXED, "relaxed"), (O.ITER_STRICT, "strict")):
        streams = []
        for i in range(4):
            _, eids, ds, _ = o.scan_stream(q[i], ef_search=ef, iterative=mode, max_scan_tuples=100, max_emit=150)
            pad = np.full(150, -1, np.int32)
            pad[:len(eids)] = eids
            streams.append(pad)
        out["iter_" + tag] = np.stack(streams)
    return out


def main():
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **generate(name))
        print("wrote", name)
    for name in LIVE_CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **generate_live(name))
        print("wrote", name)


if __name__ == "__main__":
    main()
