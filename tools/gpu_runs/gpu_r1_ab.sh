mkdir -p gpurun_out
( for v in pf5 pf4; do
echo "== variant $v: first quad of the next segment prefetched"; MOC_B200_LIB=$PWD/simplemoc_b200/variants/libmoc_$v.so python tools/probe.py default 2>&1 | tail -5 | head -4
done
echo "== default lib"; python tools/probe.py default 2>&1 | tail -5 | head -4
MOC_B200_LIB=$PWD/simplemoc_b200/variants/libmoc_pf4.so python -m pytest tests -m gpu -q -x -k "sweep_table_mode or lane_mappings or sfu" 2>&1 | tail -3 ) | tee gpurun_out/variants_ab.log
