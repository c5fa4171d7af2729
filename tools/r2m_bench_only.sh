set -u
N=${1:-2}
out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 --no-recall-curve --build-rows 0 > $out/r2m_bench_n$N.json 2> $out/r2m_bench_n$N.err; echo "bench rc=$?"
tail -3 $out/r2m_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2m_bench_n$N.json').read().strip().splitlines()[-1])
print('scan', d['value'], d['e2e']['value'])
print('flat', json.dumps(d['flat_knn'])[:1200])
print('strong', json.dumps(d.get('strong_scaling_scan')))
PY
