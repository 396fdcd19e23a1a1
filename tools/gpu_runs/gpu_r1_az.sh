mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_az.log
( timeout 900 python tools/fuzz_parity.py 90 77 2>&1 | grep -i "mismatch\|cases\|Traceback\|Error" | tail -5 ) | tee gpurun_out/fuzz_az.log
python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident ms', round(d['ms_per_step'],1), 'sweep', round(d['sweep_ms'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'e2e value %.4e' % d['e2e']['value'])" | tee gpurun_out/e2e_az.log
python bench.py --workload small --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('small: resident ms', round(d['ms_per_step'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'e2e value %.4e' % d['e2e']['value'])" | tee -a gpurun_out/e2e_az.log
