#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*" | tee -a gpurun_out/aba_recpolicy2.log; env "$@" timeout 300 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff --skip-generic --algos aba 2>&1 | grep -E "generated|rror" | tee -a gpurun_out/aba_recpolicy2.log; }
run BRBD_GEN_REC_POLICY=1
run BRBD_GEN_REC_POLICY=2
run BRBD_GEN_REC_POLICY=1 BRBD_GEN_REC_KEEP=0.75
run BRBD_GEN_REC_POLICY=2 BRBD_GEN_REC_KEEP=0.75
run BRBD_GEN_REC_POLICY=1 BRBD_GEN_REC_KEEP=0.5
timeout 300 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "specialized_kernels or bench_config" 2>&1 | tail -2
