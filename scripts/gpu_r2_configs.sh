#!/bin/bash
# C3 / C4 bench lines (BASELINE configs[2], configs[3]) and the parity suite again
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --config C3 --steps 5 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; echo "C3 rc=$?"; tail -c 300 gpurun_out/bench_C3.err
timeout 900 python bench.py --config C4 --steps 3 > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err; echo "C4 rc=$?"; tail -c 300 gpurun_out/bench_C4.err
timeout 900 python bench.py --config C4 --steps 3 --generic --no-cpu --no-e2e > gpurun_out/bench_C4_generic.json 2>> gpurun_out/bench_C4.err
