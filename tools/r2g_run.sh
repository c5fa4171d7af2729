set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_iterative_gpu.py tests/test_flat_tc_gpu.py tests/test_flat_gpu.py tests/test_insert_gpu.py tests/test_vacuum_gpu.py tests/test_vacuum_repair_gpu.py tests/test_group_gpu.py -q -x --timeout 600 > $out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/r2g_pytest.log
for n in 125000 250000 500000 1000000; do python tools/flat_probe.py $n 10000 10 vector_l2_ops 128 >> $out/r2g_flat_phases.log 2>&1; done; echo "flat probe rc=$?"
grep -A1 "n=125000\|n=250000" $out/r2g_flat_phases.log | tail -8
timeout 900 python bench.py --config c4 --n 2000000 --steps 50 --warmup 5 > $out/r2g_bench_c4_2m.json 2> $out/r2g_bench_c4_2m.err; echo "c4 2M rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_c4_2m.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('iterative_scan'),indent=1)[:3000])
PY
