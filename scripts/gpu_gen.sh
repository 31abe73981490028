#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py -q -x -k "specialized_derivative" 2>&1 | tail -3
timeout 900 python bench.py --config C3 --steps 5 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; echo "C3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_C3.json')); print('C3 value %.3e ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value'])); [print('  ',k,v['kernel'][:40],'%.3f ms'%v['ms_per_launch'], 'hbm %.3f' % v['hbm_frac']) for k,v in d['kernels'].items()]"
