#!/bin/bash
# N-GPU visit: parity tests (multi-device pool tests run when >1 GPU is visible), smoke, bench at N=1 and N=$1 launched as the driver does
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_multi.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2> gpurun_out/bench1.err | tee gpurun_out/bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/benchN.err | tee gpurun_out/bench_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>> gpurun_out/benchN.err | tee gpurun_out/bench_ref_n$N.json
tail -3 gpurun_out/benchN.err
