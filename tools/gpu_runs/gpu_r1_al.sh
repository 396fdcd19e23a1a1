mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_al.log
( timeout 1500 python tools/fuzz_parity.py 200 11 2>&1 | grep -i "mismatch\|refused\|cases\|skipped" | tail -30 ) | tee gpurun_out/fuzz_al.log
