#!/bin/bash
mkdir -p gpurun_out
CFG_S="bulk:12:64:1,bulk:6:64:2,bulk:8:96:1,bulk:4:96:2,bulk:5:128:1,bulk:6:128:1,bulk:12:32:2,bulk:24:32:1,bulk:3:224:1,bulk:4:160:1,bulk:5:64:2,bulk:8:64:1,bulk:3:96:2,bulk:10:64:1"
CFG_T="bulk:3:96:2,bulk:4:64:2,bulk:6:96:1,bulk:5:64:2,bulk:8:64:1,bulk:11:32:2,bulk:22:32:1,bulk:2:128:2,bulk:2:96:2,bulk:4:128:1,bulk:3:160:1,bulk:2:224:1"
for B in 65536 1048576; do
  R=20; [ $B -gt 100000 ] && R=5
  timeout 600 python scripts/crba_compact_sweep.py --batch $B --reps $R --configs "$CFG_S" simple_humanoid_ff 2>&1 | grep -v Warning | tee -a gpurun_out/crba_bulk2.log
  timeout 600 python scripts/crba_compact_sweep.py --batch $B --reps $R --configs "$CFG_T" talos_reduced_ff 2>&1 | grep -v Warning | tee -a gpurun_out/crba_bulk2.log
done
timeout 300 python scripts/crba_compact_sweep.py --batch 65536 --configs "bulk:1:64:2,bulk:2:64:1,bulk:6:32:1,bulk:3:64:1" manipulator humanoid_random 2>&1 | grep -v Warning | tee -a gpurun_out/crba_bulk2.log
