#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none -k regex:"brbd_gen_" -s 4 -c 2 -f -o gpurun_out/prof_C3 python bench.py --config C3 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_C3.log 2>&1; tail -1 gpurun_out/ncu_C3.log
