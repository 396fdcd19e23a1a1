mkdir -p gpurun_out
( timeout 40 python -m pytest tests -m gpu -q -k "exchange" 2>&1 | tail -3 ) | tee gpurun_out/pytest_gpu_n2_bi.log
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n2_bi.json 2> gpurun_out/bench_n2_bi.err
tail -1 gpurun_out/bench_n2_bi.json | cut -c1-330; tail -2 gpurun_out/bench_n2_bi.err | cut -c1-200
