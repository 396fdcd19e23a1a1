mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_af.json 2> gpurun_out/bench_af.err; tail -c 2500 gpurun_out/bench_af.json; tail -3 gpurun_out/bench_af.err
python bench.py --steps 3 --warmup 3 --exp sfu --no-e2e --no-cpu-baseline > gpurun_out/bench_af_sfu.json 2> gpurun_out/bench_af_sfu.err; tail -c 900 gpurun_out/bench_af_sfu.json
