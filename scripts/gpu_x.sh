#!/bin/bash
mkdir -p gpurun_out
for x in 0 8 24; do
echo "== BRBD_ABA_X=$x"
BRBD_ABA_X=$x timeout 300 python scripts/bench_all.py --models simple_humanoid_ff,talos_reduced_ff --algos aba --reps 9 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['ms'], 'ms  fp64', round(d['fp64_frac_of_measured'], 3))"
done | tee gpurun_out/aba_x.txt
