python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/probe.py default 2>&1 | tee gpurun_out/probe3_mb1.log
for mb in 6 8; do MOC_B200_LIB=$PWD/simplemoc_b200/_exp/libmoc_mb$mb.so python tools/probe.py default 2>&1 | tee gpurun_out/probe3_mb$mb.log; done
