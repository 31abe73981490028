#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff 2>&1 | grep -E "generated|rror" | tee gpurun_out/gen_quick.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"brbd_gen_aba" -s 3 -c 1 -f -o gpurun_out/prof_gen_aba \
  python scripts/gen_quick.py simple_humanoid_ff --reps 2 > gpurun_out/ncu_gen.log 2>&1
tail -2 gpurun_out/ncu_gen.log
