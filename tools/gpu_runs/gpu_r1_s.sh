mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/pytest_gpu_s.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_s.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; tail -c 3000 gpurun_out/bench_s.json; tail -5 gpurun_out/bench_s.err
