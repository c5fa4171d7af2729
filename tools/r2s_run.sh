set -u
out=gpurun_out; mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:flat_tc2_kernel -s 4 -c 2 -o $out/r2s_flat_tc2_125k python tools/flat_probe.py 125000 10000 10 vector_l2_ops 128 > $out/r2s_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $out | grep r2s
