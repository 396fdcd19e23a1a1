mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_bd.log
: > gpurun_out/walk_variants_bd.log
for v in base hash_only compact_only new; do
  echo "== $v" | tee -a gpurun_out/walk_variants_bd.log
  ( MOC_B200_LIB=simplemoc_b200/_exp/$v.so python tools/probe.py default 2>&1 | grep "kernel=" ) | tee -a gpurun_out/walk_variants_bd.log
done
for v in base new; do
  echo "== small $v" | tee -a gpurun_out/walk_variants_bd.log
  ( MOC_B200_LIB=simplemoc_b200/_exp/$v.so python tools/probe.py small 2>&1 | grep "kernel=" ) | tee -a gpurun_out/walk_variants_bd.log
done
