mkdir -p gpurun_out
for v in base coef; do
  echo "== variant $v"; MOC_B200_LIB=$PWD/simplemoc_b200/variants/libmoc_$v.so python tools/parity_stats.py 2>&1 | tail -10
done | tee gpurun_out/parity_stats_q.log
echo "== default lib" | tee -a gpurun_out/parity_stats_q.log; python tools/parity_stats.py 2>&1 | tail -10 | tee -a gpurun_out/parity_stats_q.log
