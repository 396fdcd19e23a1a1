set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_e.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_e.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['sweep_ms'], d['phases_ms'], d['gpu_launches'])"; tail -5 gpurun_out/bench_e.err
ncu --set full --clock-control none --import-source on -k regex:stack_walk -s 4 -c 4 -o gpurun_out/prof_walk_e -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_walk_e.log 2>&1
