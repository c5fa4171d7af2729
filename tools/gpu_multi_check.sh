# Multi-GPU check (run from the repository root):
#   /usr/local/graft/bin/gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi_check.sh N'
# the merge / group tests, then bench.py under torchrun on N GPUs without the single-GPU-only sections: weak-scaling scan,
# one caller batch split over the ranks, row-sharded exact kNN with its per-phase times (profiles/r2_bench_{2,8}gpu.json).
set -u
N=${1:-2}
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_flat_gpu.py tests/test_group_gpu.py -q --timeout 200 > $out/r2w_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -2 $out/r2w_pytest_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 --no-recall-curve --build-rows 0 > $out/r2w_bench_n$N.json 2> $out/r2w_bench_n$N.err; echo "bench rc=$?"
tail -2 $out/r2w_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2w_bench_n$N.json').read().strip().splitlines()[-1])
print('scan', d['value'], d['e2e']['value'])
f=d['flat_knn']; print('flat', f['value'], f['ms_per_batch'], f.get('phase_ms'), f['ids_match_cpu_oracle'], f['roofline']['frac'])
print('strong', json.dumps(d.get('strong_scaling_scan'))[:400])
PY
