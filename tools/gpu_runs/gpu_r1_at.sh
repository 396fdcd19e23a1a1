mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:stack_walk_block -s 4 -c 4 -o gpurun_out/prof_walk_block_at -f python bench.py --workload small --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_walk_block_at.log 2>&1
tail -2 gpurun_out/prof_walk_block_at.log
