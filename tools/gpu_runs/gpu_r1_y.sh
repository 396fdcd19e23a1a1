mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_y.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_y.json 2> gpurun_out/bench_n2_y.err
tail -c 1600 gpurun_out/bench_n2_y.json; tail -3 gpurun_out/bench_n2_y.err
