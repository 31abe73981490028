#!/bin/bash
mkdir -p gpurun_out
for v in tma tmem; do
echo "== BRBD_CRBA_V=$v"
BRBD_CRBA_V=$v timeout 300 python scripts/bench_all.py --models talos_reduced_ff,humanoid_random,humanoid --algos crba --reps 9 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['ms'], 'ms  hbm', round(d['hbm_frac_of_measured'], 3))"
done | tee gpurun_out/crba_x.txt
