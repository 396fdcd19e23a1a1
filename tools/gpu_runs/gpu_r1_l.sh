mkdir -p gpurun_out
for c in 8 16 32 64; do
MOC_B200_STREAM_CHUNKS=$c python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_l_$c.json 2> gpurun_out/bench_l_$c.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_l_$c.json').read().strip().splitlines()[-1]);print($c, d['sweep_ms'], d['e2e']['ms_per_step'], d['e2e']['value'])"
done
