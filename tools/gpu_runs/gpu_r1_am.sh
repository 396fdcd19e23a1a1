mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_am.log
