#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/derivs_gen_sweep.py --nts 224 2>&1 | grep -v Warning | tee gpurun_out/derivs_policy.log
timeout 300 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "derivative_kernels or c3_manipulator" 2>&1 | tail -2
