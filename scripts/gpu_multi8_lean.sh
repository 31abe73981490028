#!/bin/bash
# 8 GPUs of one box, lean: multi-device pool test, bench lines at N = 8 (C2, C4) and N = 2, 4 (C2)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_large.py -q -k "multi_device" 2>&1 | tail -2
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench N=$n rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --config C4 --steps 3 --warmup 3 > gpurun_out/bench_C4_n8.json 2> gpurun_out/bench_C4_n8.err; echo "C4 N=8 rc=$?"
python - <<'PY'
import json
for f in ("bench_n2","bench_n4","bench_n8","bench_C4_n8"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value %.3e e2e %.3e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    except Exception as e: print(f, "failed", e)
PY
