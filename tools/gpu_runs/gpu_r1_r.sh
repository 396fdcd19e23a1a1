mkdir -p gpurun_out
(for c in "tiny 11" "tiny 3" "mini104 11" "mini_default_in 11" "odd 11"; do python tools/parity_spread.py $c 12; done) 2>&1 | tee gpurun_out/parity_spread_r.log
