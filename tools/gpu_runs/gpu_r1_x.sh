mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_x.log
timeout 900 python bench.py --decomp-ax 2 --device-build --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_x.json 2> gpurun_out/bench_c5_x.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c5_x.json'));print('config5', d['value'], d['ms_per_step'], d['phases_ms'])"; tail -3 gpurun_out/bench_c5_x.err
python tools/probe.py default 2>&1 | tail -5
