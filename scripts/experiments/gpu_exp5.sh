#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "store_modes or packed or derivative_kernels" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_new.log
timeout 600 python scripts/derivs_gen_sweep.py 2>&1 | grep -v Warning | tee gpurun_out/derivs_gen_sweep.log
timeout 600 python scripts/derivs_gen_sweep.py --batch 65536 --reps 20 --nts 64,128,192 manipulator mixed 2>&1 | grep -v Warning | tee -a gpurun_out/derivs_gen_sweep.log
