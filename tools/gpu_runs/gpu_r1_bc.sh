mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_bc.log
( python bench.py 2>&1 | tail -1 ) | tee gpurun_out/bench_default_bc.json
