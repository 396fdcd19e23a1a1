mkdir -p gpurun_out
: > gpurun_out/walk_split_bh.log
for v in simplemoc_b200/_exp/prev.so simplemoc_b200/libmoc_b200.so; do
  echo "== $v" | tee -a gpurun_out/walk_split_bh.log
  ( MOC_B200_LIB=$v python tools/probe.py default 2>&1 | grep "kernel=" ) | tee -a gpurun_out/walk_split_bh.log
done
( python -m pytest tests -m gpu -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_bh.log
( timeout 300 python tools/fuzz_parity.py 40 99 2>&1 | grep -i "mismatch\|cases\|Traceback\|Error" | tail -4 ) | tee -a gpurun_out/walk_split_bh.log
( python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee gpurun_out/smoke_bh.log
( python bench.py 2>&1 | tail -1 ) | tee gpurun_out/bench_default_bh.json | cut -c1-400
