export MOC_B200_LIB=$PWD/simplemoc_b200/_exp/libmoc_pf.so
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/probe.py default 2>&1 | grep -v "^  renorm" | tee gpurun_out/probe8_pf.log
ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -o gpurun_out/prof_att5_pf -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_att5_pf.log 2>&1
