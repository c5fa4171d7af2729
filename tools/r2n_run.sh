set -u
out=gpurun_out; mkdir -p $out
for n in 125000 500000; do
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file $out/r2n_flat_launches_$n.csv python tools/flat_probe.py $n 10000 10 vector_l2_ops 128 > $out/r2n_ncu_$n.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2n_flat_launches_$n.csv')))
hdr=None
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        print(d['ID'], d['Kernel Name'][:40], d['Grid Size'] if 'Grid Size' in d else '', d['Metric Name'], d['Metric Value'])
PY
done
