mkdir -p gpurun_out
( python -m pytest tests/test_driver.py -m gpu -q 2>&1 | tail -12 ) | tee gpurun_out/pytest_gpu_ao.log
python - <<'P'
import sys; sys.path.insert(0,'tests')
from oracle_lib import CASES, write_input_file
write_input_file('/tmp/mini104.in', CASES['mini104'])
P
SMOC_SEED=4 oracle/_ref/SimpleMOC-dropin -i /tmp/mini104.in 2>&1 | tail -14 | tee gpurun_out/ref_main_on_gpu_ao.log
