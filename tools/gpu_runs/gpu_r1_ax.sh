mkdir -p gpurun_out
timeout 900 python bench.py --decomp-ax 2 --device-build --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_ax.json 2> gpurun_out/bench_c5_ax.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c5_ax.json'));print('config5', d['value'], d['ms_per_step'], d['phases_ms']); print(json.dumps(d['roofline'])[:900])"; tail -3 gpurun_out/bench_c5_ax.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('default', d['value'], d['roofline']['frac'], d['roofline']['note'])"
