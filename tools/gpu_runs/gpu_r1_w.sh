mkdir -p gpurun_out
timeout 900 python bench.py --decomp-ax 2 --device-build --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_w.json 2> gpurun_out/bench_c5_w.err; tail -c 1800 gpurun_out/bench_c5_w.json; tail -3 gpurun_out/bench_c5_w.err
timeout 300 python bench.py --workload small --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_w.json 2> gpurun_out/bench_small_w.err; tail -c 900 gpurun_out/bench_small_w.json
timeout 300 python bench.py --workload default_in --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_in_w.json 2> gpurun_out/bench_default_in_w.err; tail -c 900 gpurun_out/bench_default_in_w.json
for g in 32 64 128; do timeout 300 python bench.py --egroups $g --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_eg${g}_w.json 2> gpurun_out/bench_eg${g}_w.err; python -c "
import json;d=json.load(open('gpurun_out/bench_eg${g}_w.json'));print($g, d['value'], d['ms_per_step'], d['phases_ms'])"; done
