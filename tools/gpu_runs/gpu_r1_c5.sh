set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=memory.total,memory.used --format=csv,noheader
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --device-build --decomp-ax 2 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['sweep_ms'], d['phases_ms'], d['config'], d['roofline'])"; tail -5 gpurun_out/bench_c5.err
