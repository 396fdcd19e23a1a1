mkdir -p gpurun_out
timeout 900 python bench.py --decomp-ax 2 --device-build --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_aq.json 2> gpurun_out/bench_c5_aq.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c5_aq.json'));print('config5', d['value'], d['ms_per_step'], d['phases_ms'])"; tail -3 gpurun_out/bench_c5_aq.err
timeout 300 python bench.py --workload small --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_aq.json 2> gpurun_out/bench_small_aq.err; python -c "
import json;d=json.load(open('gpurun_out/bench_small_aq.json'));print('small', d['value'], d['ms_per_step'], d['phases_ms'], d['e2e']['value'])"
