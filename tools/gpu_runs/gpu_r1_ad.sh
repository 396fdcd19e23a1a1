# config 5 (131 GB per GPU, device-built) on 2 GPUs: exchange staging + record buffers must fit; then the exchange tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --decomp-ax 2 --device-build --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_n2_ad.json 2> gpurun_out/bench_c5_n2_ad.err
tail -c 1500 gpurun_out/bench_c5_n2_ad.json; tail -5 gpurun_out/bench_c5_n2_ad.err
( python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_ad.log
