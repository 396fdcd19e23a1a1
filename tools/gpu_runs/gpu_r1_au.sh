mkdir -p gpurun_out
( timeout 1700 python tools/fuzz_parity.py 400 2026 2>&1 | grep -i "mismatch\|refused\|cases\|Error\|Traceback" | tail -12 ) | tee gpurun_out/fuzz_au.log
