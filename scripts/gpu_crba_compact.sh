#!/bin/bash
# Generated CRBA with compact staging: sweep at 65 536 and 2^20 configurations.  Output: gpurun_out/crba_compact.log
mkdir -p gpurun_out
timeout 900 python scripts/crba_compact_sweep.py --baseline 2>&1 | grep -v Warning | tee gpurun_out/crba_compact.log
timeout 600 python scripts/crba_compact_sweep.py --baseline --batch 1048576 --reps 5 --configs "4:640,4:768,5:640,5:704,6:608,8:480" 2>&1 | grep -v Warning | tee -a gpurun_out/crba_compact.log
