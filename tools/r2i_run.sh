set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_iterative_gpu.py tests/test_flat_tc_gpu.py tests/test_flat_gpu.py tests/test_search_gpu.py tests/test_bench_configs_gpu.py -q -x --timeout 600 > $out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2i_pytest.log
for n in 125000 250000 500000 1000000; do python tools/flat_probe.py $n 10000 10 vector_l2_ops 128 >> $out/r2i_flat_phases.log 2>&1; done; echo "flat probe rc=$?"
grep -A1 "path=2" $out/r2i_flat_phases.log | grep phases | awk 'NR%4==0'
timeout 900 python bench.py --config c4 --n 2000000 --steps 50 --warmup 5 > $out/r2i_bench_c4_2m.json 2> $out/r2i_bench_c4_2m.err; echo "c4 2M rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_c4_2m.json').read().strip().splitlines()[-1])
for s in d.get('iterative_scan'): print({k:s[k] for k in ('filter','queries','value','seconds','rounds','calls','kernel_ms','streams_match_oracle')}, s['cpu_baseline']['value'])
PY
python bench.py --build-rows 0 --no-recall-curve > $out/r2i_bench_nobuild.json 2> $out/r2i_bench_nobuild.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_nobuild.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['flat_knn']['value'], d['ef_search_100']['value'], d['clocks'])
PY
