#!/bin/bash
# ncu --set full capture of the bench kernels (one launch each after warm-up) -> gpurun_out/prof_<tag>.ncu-rep
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${2:-dfs_kernel}" -s ${3:-6} -c ${4:-2} -f -o gpurun_out/prof_$TAG \
  python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/
