#!/bin/bash
# compute-sanitizer over the kernels of the second half of round 2 (scripts/sanitizer_run2.py); logs under gpurun_out/
mkdir -p gpurun_out
timeout 80 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitizer_run2.py > gpurun_out/sanitizer2_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|done|Invalid|rror" gpurun_out/sanitizer2_memcheck.log | tail -4
timeout 60 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitizer_run2.py simple_humanoid_ff > gpurun_out/sanitizer2_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|done|hazard|rror" gpurun_out/sanitizer2_racecheck.log | tail -4
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "minverse" 2>&1 | tail -1
