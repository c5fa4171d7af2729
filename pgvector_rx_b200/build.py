"""Builds libpgvrx_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m pgvector_rx_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpgvrx_b200.so")
SOURCES = ["capi.cu", "hnsw_search.cu", "hnsw_build.cu", "hnsw_team.cu", "build_host.cu", "flat_knn.cu", "flat_host.cu", "flat_tc.cu", "flat_tc_host.cu", "hnsw_iter.cu", "iter_host.cu", "gather_probe.cu", "group.cu", "seam.cu", "page_export.cpp", "page_write.cpp"]
HEADERS = ["hnsw_device.cuh", "hnsw_kernels.h", "index_internal.h", os.path.join("..", "..", "include", "pgvrx_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "--fmad=false",   # distances must not contract mul+add
]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, timing=False, cpasync_gather=False):
    """Compile every source to an object in parallel, then link libpgvrx_b200.so.

    timing=True builds libpgvrx_b200_timing.so with per-phase clock64() instrumentation (a profiling
    aid selected with PGVRX_LIB=...; never the shipped library); cpasync_gather=True builds the
    measured-slower cp.async row gather for comparison."""
    from concurrent.futures import ThreadPoolExecutor
    out = LIB.replace(".so", "_timing.so") if timing else LIB
    if cpasync_gather:
        out = out.replace(".so", "_cpasync.so")
    variant = not (out == LIB)
    if not force and not variant and not is_stale():
        return LIB
    objdir = os.path.join(HERE, "build", os.path.basename(out).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    defs = (["-DPGVRX_PHASE_TIMING"] if timing else []) + (["-DPGVRX_GATHER_CPASYNC"] if cpasync_gather else [])
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + defs
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))

    def compile_one(src):
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        sp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), hdr_t):
            return obj, 0, ""
        r = subprocess.run([nvcc_path()] + cflags + ["-c", "-o", obj, sp], capture_output=True, text=True)
        return obj, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    log = "".join(r[2] for r in res)
    if verbose or any(r[1] for r in res):
        sys.stderr.write(log)
    if any(r[1] for r in res):
        raise RuntimeError("nvcc failed compiling libpgvrx_b200.so")
    r = subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + [x[0] for x in res],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libpgvrx_b200.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, timing="--timing" in sys.argv,
                cpasync_gather="--cpasync-gather" in sys.argv))
