#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_large.py tests/test_shim.py -m gpu -q -x -k "host_threads or shim or eigen or packed" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_new.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_ht.json 2> gpurun_out/bench_ht.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_ht.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_ht.json').read().strip().splitlines()[-1]); print('e2e', d['e2e'])"
