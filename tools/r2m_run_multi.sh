# multi-GPU check: run with `gpurun --gpus N -- 'bash tools/r2m_run_multi.sh N'`
set -u
N=${1:-2}
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_flat_gpu.py tests/test_flat_tc_gpu.py -q --timeout 300 > $out/r2m_pytest_group_n$N.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2m_pytest_group_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 --no-recall-curve --build-rows 0 > $out/r2m_bench_n$N.json 2> $out/r2m_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2m_bench_n$N.json').read().strip().splitlines()[-1])
print('scan', d['value'], d['e2e']['value'])
print('flat', json.dumps(d['flat_knn'])[:900])
print('strong', json.dumps(d.get('strong_scaling_scan')))
PY
g++ -O2 -std=c++17 -pthread -Iinclude tools/pgvrx_bench.cpp -Lpgvector_rx_b200 -lpgvrx_b200 -Wl,-rpath,$PWD/pgvector_rx_b200 -o /tmp/pgvrx_bench && /tmp/pgvrx_bench --n 500000 --dim 128 --nq 10000 --steps 20 --warmup 3 --gpus $N > $out/r2m_bench_binary_n$N.json 2> $out/r2m_bench_binary_n$N.err; echo "binary rc=$?"; cat $out/r2m_bench_binary_n$N.json | head -c 1600
