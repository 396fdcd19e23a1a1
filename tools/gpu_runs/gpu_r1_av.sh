mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -k "randomised or driver" 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_av.log
( timeout 600 python tools/fuzz_parity.py 60 99 2>&1 | grep -i "mismatch\|cases\|Traceback\|Error" | tail -5 ) | tee gpurun_out/fuzz_av.log
