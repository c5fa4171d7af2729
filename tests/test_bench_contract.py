"""bench.py's JSON contract.  The reference arm (`--impl reference`: the oracle on the host cores, the one place outside
the tests where bench.py runs the checker) is executed here at a small size; the GPU arm's line is checked on the lines
recorded from B200 runs under profiles/ (the GPU arm itself refuses to run without a device)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _baseline_metric():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        return json.load(f)["metric"]


def _last_json_line(text):
    return json.loads([l for l in text.strip().splitlines() if l.startswith("{")][-1])


def test_reference_arm_line(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "2000", "--nq", "64",
                        "--steps", "2", "--warmup", "1", "--cache", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = _last_json_line(r.stdout)
    assert BASE_KEYS <= set(out), BASE_KEYS - set(out)
    assert out["impl"] == "reference" and out["metric"] == _baseline_metric() and out["unit"] == "queries/s"
    assert out["higher_is_better"] is True and out["vs_baseline"] is None and out["steps"] == 2 and out["warmup"] == 1
    assert out["value"] > 0 and abs(out["value"] - 64 / (out["ms_per_step"] / 1e3)) < 1e-6 * out["value"]
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == out["value"] and cb["sample"]
    assert out["e2e"] == {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in out["config"] and "model" not in out["config"]
    assert 0.5 < out["recall_at_10"] <= 1.0                       # 2000 rows: the graph scan finds most true neighbours


def test_reference_arm_other_ranks_exit_quietly(tmp_path):
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n", "2000",
                        "--nq", "64", "--steps", "1", "--warmup", "1", "--cache", str(tmp_path)], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CPU path" in r.stderr and r.stdout.strip() == ""


@pytest.mark.parametrize("name,n_gpus", [("r1_bench_1gpu.json", 1), ("r1_bench_2gpu.json", 2)])
def test_recorded_gpu_lines_carry_the_contract(name, n_gpus):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip("no recorded line")
    with open(path) as f:
        out = _last_json_line(f.read())
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(out)
    assert out["metric"] == _baseline_metric() and out["n_gpus"] == n_gpus and out["scaling"] == "weak"
    assert out["gpu_launches"] >= out["steps"] and out["ids_match_cpu_oracle"] is True
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(out["clocks"])
    assert not set(out["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = out["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != out["value"]
    rf = out["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert 0 < rf["frac"] < 1.0
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0


def test_rank0_only_sections_do_not_enter_process_group_barriers():
    # At N > 1 everything after `if rank == 0:` in run_b200 runs on rank 0 alone (CPU arms, ef_search = 100, the recall
    # sweep).  ScanRunner's timed loops contain process-group barriers unless its collectives are switched off first; a
    # forgotten switch deadlocks the job until NCCL's watchdog kills it (it did, once, on the first 2-GPU run of round 2).
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    body = src[src.index("def run_b200(a):"):src.index("def lowdim_curve(")]
    alone = body[body.index("    if rank == 0:\n        import oracle_ffi as O"):]
    off = alone.index("runner.collective = False")
    for call in ("runner.device_timed(", "runner.wall_timed(", "ef_sweep("):
        first = alone.find(call)
        assert first < 0 or off < first, call
    runner = src[src.index("class ScanRunner"):src.index("def parity_sample(")]
    assert runner.count("self.world > 1 and self.collective") == 3      # barrier, device_timed, wall_timed
