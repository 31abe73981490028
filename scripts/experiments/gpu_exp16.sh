#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_parity.py -m gpu -q -x -k "specialized or pool_surface or minverse or crba or stream" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_new.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
