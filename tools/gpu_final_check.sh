# One gpurun call that re-establishes a round's single-GPU evidence on a fresh B200 (run from the repository root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_final_check.sh'
# the whole GPU parity suite, smoke(), the default bench line and the reference arm; results under gpurun_out/r2v_*
# (profiles/r2_bench_1gpu_final.json, r2_bench_reference_arm.json, r2_smoke.txt, r2_pytest_gpu_tail.txt came from it).
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $out/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/r2v_smoke.log | cut -c1-400
python bench.py > $out/r2v_bench_1gpu.json 2> $out/r2v_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/r2v_bench_reference_arm.json 2> $out/r2v_bench_reference_arm.err; echo "reference arm rc=$?"; cat $out/r2v_bench_reference_arm.json | cut -c1-600
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_of_gather_ceiling'), d['parity'], d['clocks'])
print('flat', d['flat_knn']['value'], d['flat_knn']['ms_per_batch'], d['flat_knn']['roofline']['frac'], d['flat_knn'].get('candidates_per_query'))
print('build', {k:d['build'].get(k) for k in ('gpu_vectors_per_s','seconds','adjacency_identical_to_oracle')}, d['build'].get('insert',{}).get('gpu_vectors_per_s'), d['build'].get('insert',{}).get('graph_identical_to_oracle'))
print('ef100', d['ef_search_100']['value'], d['ef_search_100']['roofline']['frac'])
for c in d['recall_curve']: print(c['dataset'][:30], c.get('qps_at_recall_0.90'))
PY
