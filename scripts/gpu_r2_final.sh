#!/bin/bash
# Round-2 final evidence pass: -m gpu suite, smoke, both bench arms (C2), generic-kernel bench line, C3 / C4 lines, ncu launch list,
# ncu --set full of the step's kernels, the FP32 / FP64 sweep of configs[4].  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
timeout 600 python bench.py --generic --no-cpu --no-e2e > gpurun_out/bench_generic.json 2>> gpurun_out/bench.err
timeout 900 python bench.py --config C3 --steps 5 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; echo "C3 rc=$?"
timeout 900 python bench.py --config C4 --steps 3 > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err; echo "C4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"brbd_gen_aba|brbd_gen_crba" -s 6 -c 2 -f -o gpurun_out/prof_step \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
timeout 900 python scripts/precision_sweep.py > gpurun_out/precision_sweep.jsonl 2> gpurun_out/precision_sweep.err; echo "sweep rc=$?"
timeout 600 python scripts/bench_all.py > gpurun_out/all_algorithms.jsonl 2> gpurun_out/bench_all.err; echo "bench_all rc=$?"
ls gpurun_out/ | head -40
