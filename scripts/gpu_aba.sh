#!/bin/bash
# ABA: parity of the aba tests, then timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "aba or euler or smoke" 2>&1 | tail -3 | tee gpurun_out/pytest_aba.log
for B in 65536 524288; do
timeout 300 python scripts/bench_all.py --models simple_humanoid_ff,talos_reduced_ff,manipulator --algos aba --batch $B --reps 9 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['batch'], d['ms'], 'ms  fp64', round(d['fp64_frac_of_measured'], 3))"
done | tee gpurun_out/aba_quick.txt
