#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_large.py -q -x -k "specialized or bench_config" 2>&1 | tail -4
timeout 300 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff manipulator 2>&1 | grep -E "generated|rror" | tee gpurun_out/gen_quick.log
timeout 300 python scripts/gen_quick.py simple_humanoid_ff talos_reduced_ff --batch 1048576 --reps 5 2>&1 | grep -E "generated|rror" | tee -a gpurun_out/gen_quick.log
