#!/bin/bash
mkdir -p gpurun_out
for pf in 0 1 2; do
  echo "== BRBD_GEN_PREFETCH=$pf" | tee -a gpurun_out/gen_prefetch.log
  if [ $pf -eq 0 ]; then unset BRBD_GEN_PREFETCH; else export BRBD_GEN_PREFETCH=$pf; fi
  for rep in 1 2; do
  timeout 300 python bench.py --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: ms/step %.4f'%d['ms_per_step'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})" | tee -a gpurun_out/gen_prefetch.log
  done
done
