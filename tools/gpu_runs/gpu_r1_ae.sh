mkdir -p gpurun_out
( python -m pytest tests -m gpu -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_ae.log
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py 2>&1 | tail -25; echo "memcheck exit: $?" ) | tee gpurun_out/sanitizer_memcheck_ae.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py 2>&1 | tail -25 ) | tee gpurun_out/sanitizer_racecheck_ae.log
