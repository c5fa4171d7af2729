# ncu evidence for K1 (run from the repository root): launch list of a short bench run + one --set full capture of the scan kernel
#   /usr/local/graft/bin/gpurun --timeout 900 -- "bash tools/gpu_ncu_scan.sh"
set -u
out=gpurun_out; mkdir -p $out
short="python bench.py --steps 3 --warmup 3 --cpu-sample 160 --recall-sample 8 --build-rows 0 --no-flat --no-recall-curve --no-ef100 --no-gather-probe"
if $short > $out/r2x_short.json 2> $out/r2x_short.err; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r2x_launches.csv $short > $out/r2x_ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:hnsw_scan_kernel -s 6 -c 1 -o $out/r2x_scan_kernel $short > $out/r2x_ncu_full.log 2>&1
fi
ls -la $out | grep r2x
