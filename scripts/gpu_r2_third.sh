#!/bin/bash
# bulk-copy CRBA as the default: new tests, bench C2 / C4, launch list, ncu --set full of the step's kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "store_modes or packed or specialized or bench_config" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_new.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err
timeout 900 python bench.py --config C4 --steps 3 --no-cpu > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err; echo "C4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"brbd_gen_aba|brbd_gen_crba" -s 6 -c 2 -f -o gpurun_out/prof_step \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
