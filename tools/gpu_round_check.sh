#!/bin/bash
# One gpurun call that re-establishes the round's evidence on a fresh B200 (run from the repository root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh r2'
# 1. the GPU parity suite, 2. the bench line at ef_search 40 (the headline) and 100 (BASELINE configs[1] names both),
# 3. the reference arm, 4. the standalone C++ bench binary, 5. the ncu launch list of a short bench run and one
# --set full capture of the scan kernel (each only after the same command has exited 0 without ncu).
# Everything lands in gpurun_out/<tag>_*; copy what is to be judged into profiles/.
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench ef40 rc=$?"
python bench.py --ef 100 --build-rows 0 --no-flat > $out/${tag}_bench_1gpu_ef100.json 2> $out/${tag}_bench_1gpu_ef100.err; echo "bench ef100 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"
g++ -O2 -std=c++17 -pthread -Iinclude tools/pgvrx_bench.cpp -Lpgvector_rx_b200 -lpgvrx_b200 -Wl,-rpath,$PWD/pgvector_rx_b200 -o /tmp/pgvrx_bench \
  && /tmp/pgvrx_bench --n 1000000 --dim 128 --nq 10000 --steps 50 --warmup 5 > $out/${tag}_bench_binary_1m.json 2> $out/${tag}_bench_binary_1m.err; echo "bench binary rc=$?"
short="python bench.py --steps 3 --warmup 3 --cpu-sample 160 --recall-sample 8 --build-rows 0 --no-flat"
if $short > $out/${tag}_short.json 2> $out/${tag}_short.err; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $short > $out/${tag}_ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:hnsw_scan_kernel -s 6 -c 1 -o $out/${tag}_scan_kernel $short > $out/${tag}_ncu_full.log 2>&1
fi
ls -la $out | tail -20
