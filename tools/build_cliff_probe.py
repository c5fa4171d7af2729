"""Times the build where select_neighbors' pairwise distances no longer fit the GPU mask kernel (ef_construction > 128,
or rows too wide for 128 of them to share a CTA's shared memory) and fall back to host distances:
python tools/build_cliff_probe.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import pgvector_rx_b200 as P  # noqa: E402
from pgvector_rx_b200 import datasets as D  # noqa: E402

for name, n, dim, m, efc in (("efc 64 (masks on the GPU)", 20000, 128, 16, 64), ("efc 200 (> 128 candidates per list)", 20000, 128, 16, 200),
                             ("768-d fp32 rows, 3072 B (64 rows per group)", 8000, 768, 16, 64)):
    rows = D.gaussian_f32(n, dim, 2)
    lv = D.hnsw_levels(n, m, 4)
    ix = P.HnswIndex("vector_l2_ops", dim, m=m, ef_construction=efc, capacity=n)
    t0 = time.time()
    st = ix.build(rows, lv)
    dt = time.time() - t0
    print(json.dumps(dict(case=name, rows=n, dim=dim, m=m, efc=efc, vectors_per_s=n / dt, seconds=dt, rounds=st["n_rounds"],
                          host_distance_evals=st["n_host_dist"], pairwise_on_gpu=st["n_pair_gpu"], kernel_ms=st["kernel_ms"])), flush=True)
    ix.close()
