mkdir -p gpurun_out
true
