#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 512" "2 384" "3 256" "4 192" "5 160" "7 96"; do
  set -- $cfg
  echo "== CRBA group K=$1 NT=$2"
  BRBD_CRBA_V=gen BRBD_GEN_CRBA_K=$1 BRBD_GEN_CRBA_NT=$2 timeout 300 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff 2>&1 | grep -E "generated crba|rror" | tee -a gpurun_out/gen_quick.log
done
BRBD_CRBA_V=gen BRBD_GEN_CRBA_K=3 BRBD_GEN_CRBA_NT=256 timeout 300 python scripts/gen_quick.py simple_humanoid_ff --batch 1048576 --reps 5 2>&1 | grep -E "generated crba|rror"
