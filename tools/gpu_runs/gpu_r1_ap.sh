mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_ap.log
( timeout 1500 python tools/fuzz_parity.py 120 33 2>&1 | grep -i "mismatch\|refused\|cases\|Error\|Traceback" | tail -20 ) | tee gpurun_out/fuzz_ap.log
( python tools/probe.py small 2>&1 | tail -6 ) | tee gpurun_out/probe_small_ap.log
