#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for w in 4 5 6 7 8; do
  echo "== BRBD_ABA_WARPS=$w"
  BRBD_ABA_WARPS=$w timeout 300 python scripts/bench_all.py --models simple_humanoid_ff,talos_reduced_ff,manipulator --algos aba --reps 9 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['ms'], 'ms  fp64', round(d['fp64_frac_of_measured'], 3))"
done | tee gpurun_out/aba_sweep.txt
echo "== v3"; BRBD_ABA_V3=1 timeout 300 python scripts/bench_all.py --models simple_humanoid_ff,talos_reduced_ff,manipulator --algos aba --reps 9 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['ms'], 'ms  fp64', round(d['fp64_frac_of_measured'], 3))" | tee -a gpurun_out/aba_sweep.txt
timeout 600 python scripts/large_batch_check.py --quick 2>&1 | grep '"aba"\|euler' | cut -c1-200
