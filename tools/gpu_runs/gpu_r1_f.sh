set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_f.log
