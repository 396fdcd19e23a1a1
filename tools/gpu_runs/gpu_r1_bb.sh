mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_bb.log
( python tools/probe.py small 2>&1 | tail -1; python tools/probe.py default 2>&1 | tail -1 ) | tee gpurun_out/update_bb.log
( timeout 900 python tools/fuzz_parity.py 60 123 2>&1 | grep -i "mismatch\|cases\|Traceback\|Error" | tail -4 ) | tee -a gpurun_out/update_bb.log
