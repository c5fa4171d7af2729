"""Times the exact flat kNN (K4) on a C2-shaped table: python tools/flat_probe.py [n] [nq] [k]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgvector_rx_b200 as P
from pgvector_rx_b200 import datasets as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
opclass = sys.argv[4] if len(sys.argv) > 4 else "vector_l2_ops"
dim = int(sys.argv[5]) if len(sys.argv) > 5 else 128
elem = P.OPCLASSES[opclass][0]
if elem == P.ELEM_BIT:
    rows, q = D.random_bits(n, dim, 8), D.random_bits(nq, dim, 9)
elif elem == P.ELEM_F16:
    rows, q = D.gaussian_f32(n, dim, 5).astype(np.float16), D.gaussian_f32(nq, dim, 6).astype(np.float16)
else:
    rows, q = D.gaussian_f32(n, dim, 2), D.gaussian_f32(nq, dim, 3)
ix = P.HnswIndex(opclass.replace("cosine", "ip"), dim, capacity=n)
ix.load_rows(rows)
dev = torch.device("cuda", 0)
d_q = torch.from_numpy(q.view(np.uint8) if q.dtype != np.float32 else q).to(dev)
ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
ds = torch.empty((nq, k), dtype=torch.float32, device=dev)
for it in range(3):
    st = P.FlatStats()
    ix.flat_topk_device(d_q.data_ptr(), nq, k, ids.data_ptr(), ds.data_ptr(), stats=st)
    rb = ix.L.pgvrx_row_bytes(ix.elem, dim)
    print("flat %s n=%d nq=%d k=%d dim=%d: %.2f ms, %.1f kQPS, %.2f Tpairs-elem/s, base traffic if streamed once per q-tile %.1f GB, splits %d path %d"
          % (opclass, n, nq, k, dim, st.kernel_ms, nq / st.kernel_ms, n * nq * dim / st.kernel_ms / 1e9,
             (nq + 31) // 32 * n * rb / 1e9, st.n_splits, st.path), flush=True)
