"""Times the flat kNN on a C2-shaped table: python tools/flat_probe.py [n] [nq] [k] [opclass] [dim] [path]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
opclass = sys.argv[4] if len(sys.argv) > 4 else "vector_l2_ops"
dim = int(sys.argv[5]) if len(sys.argv) > 5 else 128
if len(sys.argv) > 6:
    os.environ["PGVRX_FLAT_PATH"] = sys.argv[6]
import pgvector_rx_b200 as P
from pgvector_rx_b200 import datasets as D

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
for it in range(4):
    st = P.FlatStats()
    ix.flat_topk_device(d_q.data_ptr(), nq, k, ids.data_ptr(), ds.data_ptr(), stats=st)
    flops = 2.0 * n * nq * dim * (3 if st.path >= 1 else 1)
    print("flat %s n=%d nq=%d k=%d dim=%d path=%d: %.2f ms, %.1f kQPS, %.1f TFLOP/s issued, splits %d, candidates/query %d, uncertified %d"
          % (opclass, n, nq, k, dim, st.path, st.kernel_ms, nq / st.kernel_ms, flops / st.kernel_ms / 1e9, st.n_splits,
             st.n_candidates // max(nq, 1), st.n_uncertified), flush=True)
    if st.path >= 1:
        print("    phases (ms): " + ", ".join("%s %.3f" % kv for kv in st.as_dict()["phase_ms"].items()), flush=True)
# the same call without stats (asynchronous), as the sharded path issues it: wall time per call, 10 calls back to back
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(10):
    ix.flat_topk_device(d_q.data_ptr(), nq, k, ids.data_ptr(), ds.data_ptr())
torch.cuda.synchronize()
print("    10 calls without stats: %.3f ms per call (wall)" % ((time.perf_counter() - t0) * 100), flush=True)
if len(sys.argv) > 7:
    # cross-check the two paths against each other on a sample
    os.environ["PGVRX_FLAT_PATH"] = "exact"
    i2 = torch.empty((256, k), dtype=torch.int32, device=dev)
    d2 = torch.empty((256, k), dtype=torch.float32, device=dev)
    ix.flat_topk_device(d_q.data_ptr(), 256, k, i2.data_ptr(), d2.data_ptr())
    torch.cuda.synchronize()
    print("paths agree on 256 queries:", bool((i2 == ids[:256]).all().item()) and bool((d2 == ds[:256]).all().item()))
