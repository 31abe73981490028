#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "store_modes or packed" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_new.log
for s in 0 50 150 400 1200; do
  echo "== BRBD_GEN_SYNC=$s" | tee -a gpurun_out/gen_sync.log
  BRBD_GEN_SYNC=$s timeout 300 python scripts/gen_quick.py simple_humanoid_ff --skip-generic --algos rnea,aba 2>&1 | grep -E "generated|rror" | tee -a gpurun_out/gen_sync.log
  BRBD_GEN_SYNC=$s timeout 300 python scripts/gen_quick.py simple_humanoid_ff --skip-generic --algos rnea,aba --batch 1048576 --reps 5 2>&1 | grep -E "generated|rror" | tee -a gpurun_out/gen_sync.log
done
