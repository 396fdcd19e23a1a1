set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_h.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_h.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['sweep_ms'], d['phases_ms'], d['gpu_launches'], d['config']['segments_per_sweep_per_gpu'])"; tail -5 gpurun_out/bench_h.err
ncu --set full --clock-control none --import-source on -k regex:"stack_walk|attenuate" -s 5 -c 5 -o gpurun_out/prof_h -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_h.log 2>&1
