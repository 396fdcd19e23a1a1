mkdir -p gpurun_out
( for c in 16 32 64 128; do
echo "== MOC_B200_STREAM_CHUNKS=$c"; MOC_B200_STREAM_CHUNKS=$c python bench.py --steps 2 --warmup 2 --e2e-steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident ms', round(d['ms_per_step'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'e2e value %.4e' % d['e2e']['value'])"
done ) | tee gpurun_out/chunks_ac.log
