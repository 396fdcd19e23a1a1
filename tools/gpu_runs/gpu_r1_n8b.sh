# 8 GPUs: default problem 2x2x2 (config 3), config 5 (131 GB per GPU) 2x2x2, default 2x2x1 on 4, exchange tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8_b.json 2> gpurun_out/bench_n8_b.err
tail -c 700 gpurun_out/bench_n8_b.json; tail -2 gpurun_out/bench_n8_b.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --decomp-ax 2 --device-build --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_n8_b.json 2> gpurun_out/bench_c5_n8_b.err
tail -c 700 gpurun_out/bench_c5_n8_b.json; tail -2 gpurun_out/bench_c5_n8_b.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4_b.json 2> gpurun_out/bench_n4_b.err
tail -c 400 gpurun_out/bench_n4_b.json
( timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_n8b.log
