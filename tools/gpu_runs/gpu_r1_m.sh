mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "edge_cases" 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_m.log
