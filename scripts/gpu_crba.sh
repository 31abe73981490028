#!/bin/bash
# CRBA: warps-per-SM sweep (BRBD_CRBA_WARPS caps the resident warps) at batch $B (default 65536)
mkdir -p gpurun_out
for w in 5 6 7 8; do
  echo "== BRBD_CRBA_WARPS=$w batch ${B:-65536}"
  BRBD_CRBA_WARPS=$w timeout 300 python scripts/bench_all.py --models simple_humanoid_ff,talos_reduced_ff,humanoid_random --algos crba --batch ${B:-65536} --reps 7 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['ms'], 'ms  hbm', round(d['hbm_frac_of_measured'], 3))"
done | tee gpurun_out/crba_sweep_${B:-65536}.txt
