#!/bin/bash
mkdir -p gpurun_out
for nt in 256 320 384 448 512; do
  echo "== DIRECT NT=$nt"
  BRBD_GEN_DIRECT=1 BRBD_GEN_NT=$nt timeout 300 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff 2>&1 | grep -E "generated|rror" 
done 2>&1 | tee gpurun_out/gen_quick.log
