mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:stack_walk_warp -s 4 -c 4 -o gpurun_out/prof_walk_ag -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_walk_ag.log 2>&1
tail -3 gpurun_out/prof_walk_ag.log; ls -la gpurun_out/prof_walk_ag.ncu-rep
