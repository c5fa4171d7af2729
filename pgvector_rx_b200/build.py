"""Builds libpgvrx_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m pgvector_rx_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpgvrx_b200.so")
SOURCES = ["capi.cu", "hnsw_search.cu", "page_export.cpp"]
HEADERS = ["hnsw_device.cuh", "hnsw_kernels.h", os.path.join("..", "..", "include", "pgvrx_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=false",   # distances must not contract mul+add
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
    """timing=True builds libpgvrx_b200_timing.so with per-phase clock64() instrumentation
    (a profiling aid selected with PGVRX_LIB=...; never the shipped library)."""
    out = LIB.replace(".so", "_timing.so") if timing else LIB
    if cpasync_gather:
        out = out.replace(".so", "_cpasync.so")
    if not force and not timing and not cpasync_gather and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          (["-DPGVRX_PHASE_TIMING"] if timing else []) + (["-DPGVRX_GATHER_CPASYNC"] if cpasync_gather else []) + ["-o", out] + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libpgvrx_b200.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, timing="--timing" in sys.argv,
                cpasync_gather="--cpasync-gather" in sys.argv))
