#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./scripts/micro/bulk_store_bw 262144 1444 2>&1 | tee gpurun_out/bulk_store_bw.log
timeout 300 ./scripts/micro/bulk_store_bw 262144 1225 2>&1 | grep -E "mode 1|mode 0.*warps  8 depth 2" | tee -a gpurun_out/bulk_store_bw.log
for st in 0 10000 28000 60000; do
  echo "== stagger $st" | tee -a gpurun_out/stagger.log
  if [ $st -gt 0 ]; then export BRBD_GEN_CRBA_STAGGER=$st; fi
  timeout 300 python scripts/crba_compact_sweep.py --batch 1048576 --reps 5 --configs "8:480,6:512" simple_humanoid_ff 2>&1 | grep -v Warning | tee -a gpurun_out/stagger.log
done
