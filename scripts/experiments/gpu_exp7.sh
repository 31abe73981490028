#!/bin/bash
mkdir -p gpurun_out
for a in 0 1 2 3 4; do
  echo "== BRBD_GEN_ABA_P3AHEAD=$a" | tee -a gpurun_out/aba_p3ahead.log
  BRBD_GEN_ABA_P3AHEAD=$a timeout 300 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff --skip-generic --algos aba 2>&1 | grep -E "generated|rror" | tee -a gpurun_out/aba_p3ahead.log
done
timeout 300 python -m pytest tests/test_gpu_large.py tests/test_shim.py -m gpu -q -x -k "minverse or minv or shim or eigen" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_new.log
