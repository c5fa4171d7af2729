"""tools/pgvrx_bench.cpp: the standalone host-side bench binary over the C ABI (SURVEY 8(b) "what calls it": the C++
stand-in for the north star's Rust bench binary).  It links against libpgvrx_b200.so only -- no Python, no torch.
CPU: it compiles warning-free and, without a device, refuses to run (no CPU fallback).  GPU: CREATE INDEX, replicas
through export -> load, one scanning host thread per replica, recall against the exact flat scan."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pgvector_rx_b200")


def compile_bench(tmp_path):
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "pgvrx_bench")
    subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tools", "pgvrx_bench.cpp"), "-L", LIBDIR, "-lpgvrx_b200",
                    "-Wl,-rpath," + LIBDIR, "-o", exe], check=True, capture_output=True, text=True)
    return exe


def _has_device():
    import pgvector_rx_b200 as P
    return P.lib().pgvrx_device_count() > 0


def test_bench_binary_compiles_and_refuses_to_run_without_a_device(tmp_path):
    exe = compile_bench(tmp_path)
    r = subprocess.run([exe, "--bogus"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "usage:" in r.stderr
    if _has_device():
        pytest.skip("a device is present: the refusal path cannot be exercised")
    r = subprocess.run([exe, "--n", "1000", "--nq", "10"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3, r.stderr
    assert "no CPU path" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_bench_binary_builds_replicates_and_scans(tmp_path):
    exe = compile_bench(tmp_path)
    # two replicas on whatever devices exist (--oversubscribe folds them onto one GPU): the second handle is loaded
    # from the first one's exported graph and scanned by its own host thread
    r = subprocess.run([exe, "--n", "20000", "--dim", "32", "--nq", "2000", "--steps", "3", "--warmup", "1", "--gpus", "2",
                        "--oversubscribe", "--recall-queries", "500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["n_gpus"] == 2 and out["steps"] == 3
    assert out["value"] > 0 and out["kernel_qps"] > 0
    assert out["gpu_launches"] >= 2 * 3                       # at least one kernel per call: the CUDA path ran
    assert out["full_result_sets"] == 500                     # k rows per query (k <= ef_search)
    assert out["recall_at_k"] > 0.8                           # 32-d N(0,1), 20k rows, ef_search 40
    assert out["n_dist"] > 0 and out["build"]["vectors_per_s"] > 0
    # the same devices behind one pgvrx_group: row-sharded exact kNN (merge over the members' lists) and one batch split
    # by queries, both checked inside the binary against the single-device calls
    grp = out["group"]
    assert grp["members"] == 2 and grp["flat_equals_single_table"] == 1 and grp["scan_equals_single_device"] == 1
    assert grp["flat_sharded_qps"] > 0 and grp["scan_one_batch_qps"] > 0 and grp["lists_read_in_place"] >= 1
