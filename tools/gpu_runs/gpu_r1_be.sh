mkdir -p gpurun_out
: > gpurun_out/diag_be.log
for v in base new compact_only hash_only; do
  ( MOC_B200_LIB=simplemoc_b200/_exp/$v.so python tools/diag_case.py tiny_flat 11 4 2>&1 | tail -5 ) | tee -a gpurun_out/diag_be.log
done
( MOC_B200_LIB=simplemoc_b200/_exp/base.so python tools/diag_case.py tiny 11 2 2>&1 | tail -2 ) | tee -a gpurun_out/diag_be.log
( MOC_B200_LIB=simplemoc_b200/_exp/new.so python tools/diag_case.py tiny 11 2 2>&1 | tail -2 ) | tee -a gpurun_out/diag_be.log
