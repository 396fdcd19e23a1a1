# final state of round 1: gpu tests, smoke, default bench (both arms)
mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; tail -c 400 gpurun_out/bench_final_reference.json
