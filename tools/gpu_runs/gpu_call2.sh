set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
python bench.py --steps 3 --warmup 3 --exp sfu --no-e2e --no-cpu-baseline > gpurun_out/bench_default_sfu.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -o gpurun_out/prof_att -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_att.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stack_walk -s 2 -c 2 -o gpurun_out/prof_walk -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_walk.log 2>&1
ls -la gpurun_out
