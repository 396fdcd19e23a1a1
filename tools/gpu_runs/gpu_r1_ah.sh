mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -x 2>&1 | tail -5 ) | tee gpurun_out/pytest_gpu_ah.log
( python tools/probe.py default 2>&1 | tail -5 | head -4 ) | tee gpurun_out/tailplane_ah.log
