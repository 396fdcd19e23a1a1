mkdir -p gpurun_out
for rep in 1 2; do
echo "== default lib (rep $rep)"; python tools/probe.py default 2>&1 | tail -5 | head -4
echo "== variant ldg1: gathers with L1::no_allocate (rep $rep)"; MOC_B200_LIB=$PWD/simplemoc_b200/variants/libmoc_ldg1.so python tools/probe.py default 2>&1 | tail -5 | head -4
done | tee gpurun_out/variants_z.log
