mkdir -p gpurun_out
( timeout 1200 python tools/fuzz_parity.py 60 7 2>&1 | tail -70 ) | tee gpurun_out/fuzz_ai.log
