set -u
out=gpurun_out; mkdir -p $out
for mt in 256 100000; do
echo "== PGVRX_TC_PREPASS_MAX_TILES=$mt"
PGVRX_TC_PREPASS_MAX_TILES=$mt python tools/flat_probe.py 1000000 10000 10 vector_l2_ops 128 2>&1 | tail -4
PGVRX_TC_PREPASS_MAX_TILES=$mt python tools/flat_probe.py 4000000 10000 10 vector_ip_ops 96 tc check 2>&1 | tail -4
done
