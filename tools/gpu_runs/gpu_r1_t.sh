mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -x -k "sweep_table_mode or sfu_mode or track_file" 2>&1 | tail -5 ) | tee gpurun_out/pytest_gpu_t.log
for v in mb6 mb4 mb3; do
  echo "== variant $v"; MOC_B200_LIB=$PWD/simplemoc_b200/variants/libmoc_$v.so python tools/probe.py default 2>&1 | tail -5 | head -4
done | tee gpurun_out/variants_t.log
echo "== default lib (mb5, scalar table interpolation, no D)" | tee -a gpurun_out/variants_t.log; python tools/probe.py default 2>&1 | tail -5 | head -4 | tee -a gpurun_out/variants_t.log
