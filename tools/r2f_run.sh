set -u
out=gpurun_out; mkdir -p $out
python bench.py > $out/r2f_bench_1gpu.json 2> $out/r2f_bench_1gpu.err; echo "bench rc=$?"
for n in 125000 250000 500000 1000000; do python tools/flat_probe.py $n 10000 10 vector_l2_ops 128 >> $out/r2f_flat_phases.log 2>&1; done; echo "flat probe rc=$?"
short="python bench.py --steps 3 --warmup 3 --cpu-sample 160 --recall-sample 8 --build-rows 0 --no-flat --no-recall-curve --no-ef100 --no-gather-probe"
if $short > $out/r2f_short.json 2> $out/r2f_short.err; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r2f_launches.csv $short > $out/r2f_ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:hnsw_scan_kernel -s 6 -c 1 -o $out/r2f_scan_kernel $short > $out/r2f_ncu_full.log 2>&1
fi
tail -3 $out/r2f_flat_phases.log; ls -la $out | grep r2f
