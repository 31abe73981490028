#!/bin/bash
# ncu --set full of the kernels of the C3 and C4 steps (one launch each)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"brbd_gen_" -s 4 -c 2 -f -o gpurun_out/prof_C3 python bench.py --config C3 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_C3.log 2>&1; tail -1 gpurun_out/ncu_C3.log
timeout 600 ncu --set full --clock-control none -k regex:"brbd_gen_" -s 6 -c 3 -f -o gpurun_out/prof_C4 python bench.py --config C4 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_C4.log 2>&1; tail -1 gpurun_out/ncu_C4.log
