mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_ba.log
( for p in 1 0; do
echo "== MOC_B200_NO_L2_PERSIST=$p"
MOC_B200_NO_L2_PERSIST=$p python tools/probe.py default 2>&1 | tail -5 | head -4
MOC_B200_NO_L2_PERSIST=$p python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: resident ms', round(d['ms_per_step'],1), 'sweep', round(d['sweep_ms'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1))"
done ) | tee gpurun_out/l2persist_ba.log
