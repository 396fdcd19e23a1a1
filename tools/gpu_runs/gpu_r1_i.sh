set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_i_default.json 2> gpurun_out/bench_i_default.err; tail -3 gpurun_out/bench_i_default.err
python bench.py --workload small --steps 3 --warmup 3 > gpurun_out/bench_i_small.json 2> gpurun_out/bench_i_small.err; tail -3 gpurun_out/bench_i_small.err
python bench.py --workload default_in --steps 3 --warmup 3 > gpurun_out/bench_i_default_in.json 2> gpurun_out/bench_i_default_in.err; tail -3 gpurun_out/bench_i_default_in.err
python - <<'PY'
import json
for w in ("default","small","default_in"):
    d=json.loads(open(f'gpurun_out/bench_i_{w}.json').read().strip().splitlines()[-1])
    print(w, f"{d['value']:.4g}", d['ms_per_step'], d['sweep_ms'], d['phases_ms'], "e2e", d['e2e'] and (f"{d['e2e']['value']:.4g}", d['e2e']['ms_per_step']), "cpu", d['cpu_baseline'] and f"{d['cpu_baseline']['value']:.4g}", d['roofline']['l2'].get('frac_of_probe'))
PY
