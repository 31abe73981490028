#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/crba_host_sweep.py 2>&1 | grep -v Warning | tee gpurun_out/crba_host_sweep.log
