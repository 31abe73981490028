#!/bin/bash
# One GPU visit for the derivative kernels: parity tests, all-algorithm table, ncu --set full of one ∂ABA / ∂RNEA launch.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_all.py > gpurun_out/bench_all.jsonl 2> gpurun_out/bench_all.err
tail -3 gpurun_out/bench_all.err
grep -h "derivatives" gpurun_out/bench_all.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['model'], d['algo'], d['ms'], 'ms', round(d['fp64_frac_of_measured'], 3), round(d['hbm_frac_of_measured'], 3))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"derivatives_coop" -s 3 -c 1 -f -o gpurun_out/prof_dABA \
  python scripts/bench_all.py --models simple_humanoid_ff --algos aba_derivatives --reps 1 > gpurun_out/ncu_dABA.log 2>&1
tail -2 gpurun_out/ncu_dABA.log
ls -la gpurun_out/
