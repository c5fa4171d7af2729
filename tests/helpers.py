"""Shared fixtures: seeded workloads, oracle-built graphs, and the GPU index loaded with the same graph."""
import numpy as np

import oracle_ffi as O
import pgvector_rx_b200 as P
from pgvector_rx_b200 import datasets as D

ELEM_OF = {"f32": O.ELEM_F32, "f16": O.ELEM_F16, "bit": O.ELEM_BIT}


def make_rows(kind, n, dim, seed):
    """Raw element rows for a type: f32 gaussians, f16 (via the reference's encoder) or packed bits."""
    if kind == "f32":
        return D.gaussian_f32(n, dim, seed)
    if kind == "f16":
        x = D.gaussian_f32(n, dim, seed)
        L = O.lib()
        flat = np.array([L.orc_f32_to_half(float(v)) for v in x.ravel()], dtype=np.uint16)
        return flat.reshape(n, dim)
    return D.random_bits(n, dim, seed)


def normalize_rows(elem, dim, rows):
    """a16 as the build / scan callers apply it: drop zero-norm rows, unit-normalise the rest."""
    out = []
    for r in rows:
        norm, u = O.normalize(elem, dim, r)
        if norm == 0.0:
            continue
        out.append(u)
    return np.stack(out)


def oracle_index(opclass, dim, rows, m=16, efc=64, level_seed=4, tie=O.TIE_TOTAL_ORDER, levels=None):
    elem, metric, norm = P.OPCLASSES[opclass]
    if norm:
        rows = normalize_rows(elem, dim, rows)
    o = O.Oracle(elem, metric, dim, m=m, ef_construction=efc, tie_mode=tie)
    lv = D.hnsw_levels(len(rows), m, level_seed) if levels is None else levels
    o.build(rows, lv)
    return o


def vacuumed_oracle(opclass, kind, n, dim, m, efc, seed=9, frac=0.05, delete_entry=False):
    """An oracle index in the states a vacuum leaves behind: ~frac of the elements marked deleted
    (vacuum.rs:717-758), ~frac with every heap TID removed but still in the graph (vacuum.rs:118-190) and a
    few stale neighbour tuples (scan.rs:262).  Graph repair (vacuum.rs:288-640) is deliberately not run, so
    live lists still point at deleted elements -- the case scan.rs:179 exists for."""
    o = oracle_index(opclass, dim, make_rows(kind, n, dim, seed), m=m, efc=efc)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(o.n)
    k = max(1, int(o.n * frac))
    deleted, no_tid, stale = perm[:k], perm[k:2 * k], perm[2 * k:2 * k + max(1, k // 3)]
    if not delete_entry:
        deleted = deleted[deleted != o.entry]
    elif o.entry not in deleted:
        deleted = np.append(deleted, o.entry)
    o.vacuum_mark(deleted, O.VACUUM_MARK_DELETED)
    o.vacuum_mark(no_tid, O.VACUUM_REMOVE_HEAPTIDS)
    o.vacuum_mark(stale, O.VACUUM_STALE_NEIGHBORS)
    return o, deleted, no_tid, stale


def gpu_index(opclass, dim, graph, m=16, efc=64):
    ix = P.HnswIndex(opclass, dim, m=m, ef_construction=efc, capacity=max(int(graph["n"]), 1))
    ix.load_graph(graph["rows"], graph["level"], graph["adj0"], graph["upper_base"], graph["adj_up"], graph["entry"])
    return ix


def assert_same_results(gpu, orc, what=""):
    gi, gd, gc = gpu[:3]
    oi, od, oc = orc[:3]
    assert np.array_equal(gc, oc), "%s: result counts differ" % what
    bad = np.nonzero((gi != oi).any(axis=1))[0]
    assert len(bad) == 0, "%s: ids differ for %d queries, first %d:\n gpu %s\n orc %s\n gpu d %s\n orc d %s" % (
        what, len(bad), bad[0], gi[bad[0]], oi[bad[0]], gd[bad[0]], od[bad[0]])
    # distances are the reference's f32 arithmetic bit for bit (+0.0 == -0.0 allowed, as partial_cmp treats them)
    assert np.array_equal(gd, od), "%s: distances differ" % what
