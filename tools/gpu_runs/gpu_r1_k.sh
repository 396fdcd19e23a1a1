set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "device_construction or exchange or sweep_table" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_k.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-build > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_k.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['sweep_ms'], d['phases_ms'], d['config'])"; tail -5 gpurun_out/bench_k.err
