"""Freezes the on-disk format: SHA-256 of the relation the ORACLE's page writer (oracle/hnsw_oracle.cpp, following
build.rs:545-821) produces for a few seeded indexes, one per element type plus a vacuumed one.  The test
(tests/test_pages_cpu.py::test_relation_bytes_match_the_frozen_digests) makes both the oracle's writer and the product's
pgvrx_pages_write reproduce them, so that the two cannot drift together.

    python tests/golden/make_pages_golden.py        # rewrites tests/golden/pages_sha256.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import make_rows, oracle_index, vacuumed_oracle  # noqa: E402

# name -> (opclass, kind, n, dim, m, ef_construction, seed)
CASES = {
    "vector_l2": ("vector_l2_ops", "f32", 500, 24, 8, 32, 31),
    "vector_cosine_wide": ("vector_cosine_ops", "f32", 40, 1990, 16, 64, 32),     # neighbour tuples spill to the next page
    "halfvec_ip": ("halfvec_ip_ops", "f16", 400, 33, 6, 24, 33),
    "bit_jaccard": ("bit_jaccard_ops", "bit", 600, 100, 4, 16, 34),
}


def relation(name):
    opclass, kind, n, dim, m, efc, seed = CASES[name]
    return oracle_index(opclass, dim, make_rows(kind, n, dim, seed), m=m, efc=efc)


def vacuumed_relation():
    return vacuumed_oracle("vector_l2_ops", "f32", 400, 12, 8, 32, seed=35)[0]


def digest(pages):
    return {"pages": int(len(pages) // 8192), "sha256": hashlib.sha256(pages.tobytes()).hexdigest()}


if __name__ == "__main__":
    out = {name: digest(relation(name).write_pages()) for name in CASES}
    out["vector_l2_vacuumed"] = digest(vacuumed_relation().write_pages())
    with open(os.path.join(HERE, "pages_sha256.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps(out, indent=1, sort_keys=True))
