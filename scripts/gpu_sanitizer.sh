#!/bin/bash
# compute-sanitizer memcheck + racecheck over one small call of every kernel family; logs under gpurun_out/
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitizer_run.py manipulator simple_humanoid_ff mixed > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|done|Invalid|error" gpurun_out/sanitizer_memcheck.log | tail -8
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitizer_run.py manipulator simple_humanoid_ff > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|done|hazard|error" gpurun_out/sanitizer_racecheck.log | tail -8
timeout 300 python -m pytest tests/test_shim.py -q -m gpu 2>&1 | tail -3
