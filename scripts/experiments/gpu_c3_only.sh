#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config C3 --steps 5 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; echo "C3 rc=$?"; tail -c 200 gpurun_out/bench_C3.err
