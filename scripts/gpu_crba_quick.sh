#!/bin/bash
# CRBA: parity of the crba tests, then timing on three humanoids and the manipulator, both emitters
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "crba or leading or fp32 or device_pointers or error" 2>&1 | tail -4 | tee gpurun_out/pytest_crba.log
for v in tma tmem; do
echo "== BRBD_CRBA_V=$v"
BRBD_CRBA_V=$v timeout 300 python scripts/bench_all.py --models simple_humanoid_ff,talos_reduced_ff,humanoid_random,manipulator --algos crba --reps 9 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['ms'], 'ms  hbm', round(d['hbm_frac_of_measured'], 3))"
done | tee gpurun_out/crba_quick.txt
