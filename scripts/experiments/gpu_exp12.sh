#!/bin/bash
mkdir -p gpurun_out
for pol in 0 1; do
  echo "== BRBD_GEN_REC_POLICY=$pol" | tee -a gpurun_out/aba_recpolicy.log
  BRBD_GEN_REC_POLICY=$pol timeout 300 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff --skip-generic --algos aba 2>&1 | grep -E "generated|rror" | tee -a gpurun_out/aba_recpolicy.log
  BRBD_GEN_REC_POLICY=$pol timeout 300 python scripts/gen_quick.py simple_humanoid_ff --skip-generic --algos aba --batch 1048576 --reps 5 2>&1 | grep -E "generated|rror" | tee -a gpurun_out/aba_recpolicy.log
done
BRBD_GEN_REC_POLICY=1 timeout 300 python bench.py --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench with policy: ms/step', d['ms_per_step'], {k:v['ms_per_launch'] for k,v in d['kernels'].items()})" | tee -a gpurun_out/aba_recpolicy.log
