#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_parity.py -m gpu -q -x -k "minverse or minv" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_new.log
timeout 600 python scripts/minv_quick.py 2>&1 | grep -v Warning | tee gpurun_out/minv_chol.log
