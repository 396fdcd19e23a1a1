python -m pytest tests -m gpu -q 2>&1 | tail -15
ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -o gpurun_out/prof_att2 -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_att2.log 2>&1
MOC_B200_LIB=$PWD/simplemoc_b200/_exp/libmoc_mb6.so ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -o gpurun_out/prof_att2_mb6 -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_att2_mb6.log 2>&1
python tools/probe.py default 2>&1 | tee gpurun_out/probe4.log
