#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err
timeout 900 python bench.py --config C3 --steps 5 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; echo "C3 rc=$?"; tail -c 300 gpurun_out/bench_C3.err
timeout 900 python bench.py --config C4 --steps 3 > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err; echo "C4 rc=$?"; tail -c 300 gpurun_out/bench_C4.err
