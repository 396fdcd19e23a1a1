set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_b.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -c 1200 gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
