set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests -m gpu -q -k "exchange" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_n8.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print($n, f"{d['value']:.4g}", d['ms_per_step'], d['sweep_ms'], d['phases_ms'], d['config']['domains'], "e2e", d['e2e'] and (f"{d['e2e']['value']:.4g}", d['e2e']['ms_per_step']), d['keff'], d['leakage'])
except Exception as e: print("fail", e)
PY
tail -3 gpurun_out/bench_n$n.err
done
