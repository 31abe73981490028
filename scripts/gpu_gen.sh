#!/bin/bash
mkdir -p gpurun_out
for nb in 2 3 1; do
echo "== NBUF=$nb"
BRBD_GEN_CRBA_NBUF=$nb timeout 300 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff 2>&1 | grep -E "generated crba|rror" | tee -a gpurun_out/gen_quick.log
BRBD_GEN_CRBA_NBUF=$nb timeout 300 python scripts/gen_quick.py simple_humanoid_ff --batch 1048576 --reps 5 2>&1 | grep -E "generated crba|rror" | tee -a gpurun_out/gen_quick.log
done
BRBD_GEN_CRBA_NBUF=2 timeout 600 python -m pytest tests/test_gpu_large.py -q -x -k "specialized" 2>&1 | tail -3
