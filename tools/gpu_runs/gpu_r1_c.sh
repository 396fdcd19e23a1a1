set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_c.log
tools/ubench/l2_gather 2>&1 | tee gpurun_out/l2_gather.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 1500 gpurun_out/bench_c.json | head -c 700; tail -5 gpurun_out/bench_c.err
ncu --set full --clock-control none --import-source on -k regex:stack_walk -s 2 -c 2 -o gpurun_out/prof_walk_c -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_walk_c.log 2>&1
