mkdir -p gpurun_out
( timeout 1500 python tools/fuzz_parity.py 90 21 2>&1 | grep -i "mismatch\|refused\|cases\|drop-in\|Error\|Traceback" | tail -60 ) | tee gpurun_out/fuzz_an.log
