set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_build_gpu.py tests/test_insert_gpu.py tests/test_regressions_gpu.py tests/test_golden.py tests/test_vacuum_repair_gpu.py -q -x --timeout 600 > $out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2q_pytest.log
PGVRX_BUILD_DEBUG=1 python tools/build_probe.py --n 1000000 --compare data_cache/c2_1m > $out/r2q_build_c2_1m.log 2>&1; echo "build c2 rc=$?"; tail -8 $out/r2q_build_c2_1m.log | cut -c1-400
PGVRX_BUILD_DEBUG=1 PGVRX_BUILD_TEAM=0 python tools/build_probe.py --config c3 --n 20000 > $out/r2q_build_c3_20k_noteam.log 2>&1; echo "build c3 noteam rc=$?"; tail -2 $out/r2q_build_c3_20k_noteam.log | cut -c1-300
