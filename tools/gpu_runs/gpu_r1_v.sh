# final-code evidence: gpu tests, bench (both arms, sfu), launch list, DRAM traffic of a full-size K1 launch, ncu --set full of K1/K0 on a slice
mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_v.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; tail -c 1500 gpurun_out/bench_v.json; tail -3 gpurun_out/bench_v.err
python bench.py --steps 5 --warmup 3 --exp sfu --no-e2e --no-cpu-baseline > gpurun_out/bench_v_sfu.json 2> gpurun_out/bench_v_sfu.err; tail -c 600 gpurun_out/bench_v_sfu.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_v_reference.json 2> gpurun_out/bench_v_reference.err; tail -c 800 gpurun_out/bench_v_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_v.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:attenuate -s 1 -c 1 --csv --log-file gpurun_out/k1_traffic_v.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/k1_traffic_v.log 2>&1
tail -4 gpurun_out/k1_traffic_v.csv
ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -o gpurun_out/prof_att_v -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_att_v.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -o gpurun_out/prof_att_v_sfu -f python bench.py --exp sfu --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_att_v_sfu.log 2>&1
ls -la gpurun_out | tail -12
