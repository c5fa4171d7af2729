"""bench.py --config c2|c3|c4|c5 scaled down (--n / --nq): every section of every config runs on the GPU and every
comparison with the oracle it carries comes back green.  The full-size lines are recorded under profiles/ (one gpurun
call per config); these cases fail when the plumbing or the parity of any section drifts."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(tmp_path, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--cache", str(tmp_path), "--steps", "4", "--warmup", "3",
                        "--cpu-sample", "128", "--parity-sample", "128"] + list(args), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1])


def check_common(out, workload_prefix):
    assert out["config"]["workload"].startswith(workload_prefix)
    assert out["ids_match_cpu_oracle"] is True
    p = out["parity"]
    assert p["ids_identical"] and p["distances_bit_identical"] and p["counters_identical"], p
    assert out["value"] > 0 and out["e2e"]["value"] > 0 and out["e2e"]["h2d_bytes_per_step"] > 0
    rf = out["roofline"]
    assert rf["bound"] == "hbm" and rf["achieved"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["gather_ceiling"]["graph_walk"]["gb_per_s"] > 0 and rf["gather_ceiling"]["uniform"]["gb_per_s"] > 0
    assert out["cpu_baseline"]["value"] > 0 and out["cpu_baseline"]["kind"] == "port"
    assert out["flat_knn"]["ids_match_cpu_oracle"] is True
    assert 0 < out["recall_at_10"] <= 1 and abs(out["recall_at_10"] - out["recall_at_10_cpu"]) < 0.1
    assert out["clocks"]["samples"] >= 1


def test_c2_line_with_ef100_recall_curves_build_and_insert(tmp_path):
    out = run_bench(tmp_path, "--config", "c2", "--n", "30000", "--nq", "1024")
    check_common(out, "C2:")
    e = out["ef_search_100"]
    assert e["parity"]["ids_identical"] and e["parity"]["counters_identical"] and e["ids_match_cpu_oracle"]
    assert e["recall_at_10"] >= out["recall_at_10"]
    curves = out["recall_curve"]
    assert len(curves) == 2 and curves[0]["dataset"].startswith("gaussian") and curves[1]["dataset"].startswith("lowdim")
    for cv in curves:
        rec = [p["recall_at_10"] for p in cv["points"]]
        assert rec == sorted(rec) or max(abs(a - b) for a, b in zip(rec, sorted(rec))) < 0.01      # recall grows with ef_search
        assert all(p["ids_identical_on_sample"] for p in cv["points"])
    assert curves[1]["points"][-1]["recall_at_10"] > 0.9                                           # the deployable regime exists
    assert out["build"]["gpu_vectors_per_s"] > 0


def test_c3_halfvec_cosine(tmp_path):
    out = run_bench(tmp_path, "--config", "c3", "--n", "5000", "--nq", "512")
    check_common(out, "C3:")
    assert out["build"]["gpu_vectors_per_s"] > 0 and out["config"]["graph"] == "gpu-build"


def test_c4_inner_product_iterative_filtered(tmp_path):
    out = run_bench(tmp_path, "--config", "c4", "--n", "40000", "--nq", "512", "--iter-queries", "96")
    check_common(out, "C4:")
    it = out["iterative_scan"]
    assert [x["filter"] for x in it] == ["id % 50 == 0", "id % 500 == 0"]
    assert all(x["streams_match_oracle"] and x["rounds"] >= 1 for x in it)
    assert it[0]["queries_with_all_hits"] > 0
    # several refills per call under the selective filter, and a CPU arm for the section
    assert all(1 <= x["calls"] <= x["rounds"] and x["rounds"] == x["calls"] * x["refills_per_call"] for x in it)
    assert it[1]["refills_per_call"] > 1
    assert all(x["cpu_baseline"]["value"] > 0 and x["cpu_baseline"]["kind"] == "port" for x in it)


def test_c5_bit_hamming_rerank(tmp_path):
    out = run_bench(tmp_path, "--config", "c5", "--n", "40000", "--nq", "512")
    check_common(out, "C5:")
    rr = out["rerank"]
    assert rr["rerank_equals_scan_head"] is True and rr["recall_at_10_after_rerank"] >= out["recall_at_10"] - 1e-9
