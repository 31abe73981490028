#!/bin/bash
mkdir -p gpurun_out
CFG="bulk:1:256:2,bulk:1:352:2,bulk:1:224:3,bulk:1:128:2,bulk:1:512:1,bulk:2:160:2,bulk:2:128:2,bulk:2:256:1,bulk:2:352:1,bulk:3:128:2,bulk:3:256:1,bulk:3:192:1,bulk:4:96:2,bulk:4:192:1,bulk:6:128:1"
timeout 900 python scripts/crba_compact_sweep.py --configs "$CFG" 2>&1 | grep -v Warning | tee gpurun_out/crba_bulk.log
timeout 600 python scripts/crba_compact_sweep.py --batch 1048576 --reps 5 --configs "$CFG" 2>&1 | grep -v Warning | tee -a gpurun_out/crba_bulk.log
