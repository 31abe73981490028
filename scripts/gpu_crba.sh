#!/bin/bash
# CRBA experiments: parity of the crba tests, then the warps-per-SM sweep on the bench model
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "crba or smoke or sharding or position" 2>&1 | tail -5 | tee gpurun_out/pytest_crba.log
for w in 5 6 7 8; do
  echo "== BRBD_CRBA_WARPS=$w"
  BRBD_CRBA_WARPS=$w timeout 300 python scripts/bench_all.py --models simple_humanoid_ff,talos_reduced_ff,humanoid_random --algos crba --reps 9 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['ms'], 'ms  hbm', round(d['hbm_frac_of_measured'], 3))"
done | tee gpurun_out/crba_sweep.txt
