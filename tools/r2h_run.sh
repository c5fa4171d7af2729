set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_flat_tc_gpu.py tests/test_flat_gpu.py tests/test_full_size_gpu.py tests/test_group_gpu.py -q -x --timeout 600 > $out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2h_pytest.log
for n in 125000 250000 500000 1000000; do python tools/flat_probe.py $n 10000 10 vector_l2_ops 128 >> $out/r2h_flat_phases.log 2>&1; done; echo "flat probe rc=$?"
grep -A1 "path=2" $out/r2h_flat_phases.log | grep phases | awk 'NR%4==0'
python tools/k1_sweep.py --quick --no-gather --steps 100 > $out/r2h_k1_sweep.json 2> $out/r2h_k1_sweep.err; echo "sweep rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_k1_sweep.json').read().strip().splitlines()[-1])
for r in d['scan']: print(r)
PY
