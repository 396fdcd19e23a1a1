mkdir -p gpurun_out
( timeout 1200 python tools/fuzz_parity.py 26 7 2>&1 | grep -i "mismatch\|cases" ) | tee gpurun_out/fuzz_aj.log
