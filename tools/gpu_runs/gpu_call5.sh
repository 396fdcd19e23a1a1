python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python tools/probe.py default 2>&1 | grep -v "^  renorm" | tee gpurun_out/probe5_mb1.log
for mb in 4 5 6; do MOC_B200_LIB=$PWD/simplemoc_b200/_exp/libmoc_mb$mb.so python tools/probe.py default 2>&1 | grep -v "^  renorm" | tee gpurun_out/probe5_mb$mb.log; done
