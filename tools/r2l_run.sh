set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_flat_tc_gpu.py tests/test_flat_gpu.py tests/test_group_gpu.py tests/test_full_size_gpu.py -q -x --timeout 600 > $out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2l_pytest.log
for n in 125000 250000 500000 1000000; do python tools/flat_probe.py $n 10000 10 vector_l2_ops 128 >> $out/r2l_flat_phases.log 2>&1; done; echo "flat probe rc=$?"
grep -B1 "phases" $out/r2l_flat_phases.log | grep -v "^--" | awk 'NR%8==7 || NR%8==0'
python tools/flat_probe.py 125000 10000 10 vector_ip_ops 96 tc check >> $out/r2l_flat_ip.log 2>&1; tail -3 $out/r2l_flat_ip.log
python tools/flat_probe.py 200000 10000 24 vector_l2_ops 128 tc check >> $out/r2l_flat_k24.log 2>&1; tail -3 $out/r2l_flat_k24.log
