#!/bin/bash
# parity + bench only (no ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
