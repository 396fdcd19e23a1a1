mkdir -p gpurun_out
( simplemoc_b200/SimpleMOC-b200 --iters 12 2>&1 | grep -E "keff|residual|Total Time|Segments|Integrations|Last sweep|construction" | tail -32
  echo "--- host buffers, 3 iterations"
  simplemoc_b200/SimpleMOC-b200 --iters 3 --host-buffers 2>&1 | grep -E "keff|residual|Total Time|Transport Sweep Time|Integrations" | tail -10
  echo "--- default.in"
  simplemoc_b200/SimpleMOC-b200 -i /dev/stdin --iters 2 <<'IN' 2>&1 | grep -E "keff|Integrations|3D tracks" | tail -5
17
17
9
5
2
0.1
0.25
64
10
100
1
20
20
21.42
400.0
0.01
3000
0
IN
) | tee gpurun_out/driver_long_aw.log
