mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:stack_walk_warp -s 4 -c 4 -o gpurun_out/prof_walk_bg -f python bench.py --limit-tracks-2d 2000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_walk_bg.log 2>&1
tail -2 gpurun_out/prof_walk_bg.log | cut -c1-300; ls -la gpurun_out/prof_walk_bg.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default_bg.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_bg.log 2>&1
tail -1 gpurun_out/launches_bg.log | cut -c1-200; wc -l gpurun_out/launches_default_bg.csv
