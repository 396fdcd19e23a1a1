mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_bf.log
: > gpurun_out/diag_bf.log
for c in tiny_flat mini104 mini_default_in odd; do
  ( python tools/diag_case.py $c 11 3 2>&1 | tail -3 ) | tee -a gpurun_out/diag_bf.log
done
( python tools/probe.py default 2>&1 | grep "kernel=" ; python tools/probe.py small 2>&1 | grep "kernel=" ) | tee gpurun_out/walk_final_bf.log
( timeout 300 python tools/fuzz_parity.py 40 321 2>&1 | grep -i "mismatch\|cases\|Traceback\|Error" | tail -4 ) | tee -a gpurun_out/walk_final_bf.log
