set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_build_gpu.py tests/test_insert_gpu.py tests/test_vacuum_repair_gpu.py tests/test_regressions_gpu.py tests/test_seam_gpu.py -q -x --timeout 600 > $out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2j_pytest.log
PGVRX_BUILD_DEBUG=1 python tools/build_probe.py --n 1000000 --compare data_cache/c2_1m > $out/r2j_build_c2_1m.log 2>&1; echo "build c2 rc=$?"; tail -8 $out/r2j_build_c2_1m.log
PGVRX_BUILD_DEBUG=1 python tools/build_probe.py --config c3 --n 3000 > $out/r2j_build_c3_3k.log 2>&1; echo "build c3 rc=$?"; tail -3 $out/r2j_build_c3_3k.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/r2j_launches_build_c3.csv python tools/build_probe.py --config c3 --n 3000 > $out/r2j_ncu_c3.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2j_launches_build_c3.csv')))
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: agg[d['Kernel Name'][:70]].append(float(d['Metric Value'].replace(',','')))
        except: pass
for k,v in agg.items(): print("%-70s n=%5d mean=%9.1f us total=%8.2f ms" % (k,len(v),sum(v)/len(v)/1000,sum(v)/1e6))
PY
