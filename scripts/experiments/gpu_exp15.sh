#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gen_small_batch.py 2>&1 | grep -v Warning | tee gpurun_out/gen_small_batch.log
