mkdir -p gpurun_out
( python tools/probe_overlap.py 0 32 2>&1 | tail -7; python tools/probe_overlap.py 0 64 2>&1 | tail -7 ) | tee gpurun_out/overlap_g32_as.log
