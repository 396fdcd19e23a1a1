set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_j.log
cd simplemoc_b200 && ./SimpleMOC-b200 -s --iters 2 2>&1 | tail -22; cd ..
