set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_g.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -c 1300 gpurun_out/bench_g.json; tail -5 gpurun_out/bench_g.err
for g in 32 64 128; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --egroups $g > gpurun_out/bench_G$g.json 2> gpurun_out/bench_G$g.err; tail -c 300 gpurun_out/bench_G$g.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/bench_G$g.json').read().strip().splitlines()[-1]);print($g, d['value'], d['ms_per_step'], d['phases_ms'], d['roofline']['l2'])"; done
