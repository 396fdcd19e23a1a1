mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -15 ) | tee gpurun_out/pytest_gpu_u.log
python tools/probe_overlap.py 0 2>&1 | tail -9 | tee gpurun_out/overlap_u.log
python tools/probe_overlap.py 1 2>&1 | tail -9 | tee -a gpurun_out/overlap_u.log
