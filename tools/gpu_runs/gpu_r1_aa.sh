mkdir -p gpurun_out
for c in -1 0 25 50 100; do
echo "== MOC_ATT_CARVEOUT=$c"; MOC_ATT_CARVEOUT=$c python tools/probe.py default 2>&1 | tail -5 | head -4
done | tee gpurun_out/carveout_aa.log
