#!/bin/bash
# One GPU visit: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the step's kernels,
# all-algorithm table. Outputs under gpurun_out/.
bash scripts/gpu_round.sh
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"aba_rr_kernel|crba_tma_kernel" -s 6 -c 2 -f -o gpurun_out/prof_step \
  python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/ncu_step.log
timeout 900 python scripts/bench_all.py > gpurun_out/bench_all.jsonl 2> gpurun_out/bench_all.err
tail -3 gpurun_out/bench_all.err
ls -la gpurun_out/
