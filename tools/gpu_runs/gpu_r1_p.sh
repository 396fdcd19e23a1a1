mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/pytest_gpu_p.log
for v in base tail coef; do
  echo "== variant $v"; MOC_B200_LIB=$PWD/simplemoc_b200/variants/libmoc_$v.so python tools/probe.py default 2>&1 | tail -6
done | tee gpurun_out/variants_p.log
echo "== variant coef+tail (default lib)"; python tools/probe.py default 2>&1 | tail -6 | tee -a gpurun_out/variants_p.log
