set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_flat_tc_gpu.py tests/test_flat_gpu.py -q -x --timeout 600 > $out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/r2u_pytest.log
