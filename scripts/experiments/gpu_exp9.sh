#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/minv_one.py <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
os.environ["BRBD_MINV_V"] = "chol"
model = load_model("simple_humanoid_ff")
B = 65536
q, _, _ = random_inputs(model, B, 1)
tq = torch.from_numpy(np.ascontiguousarray(q.T)).cuda()
out = torch.empty((B, model.nv ** 2), dtype=torch.float64, device="cuda")
pool = pb.ModelPool(model, [0]); pool.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(3): pb.computeMinverseInParallel(1, pool, tq, out, async_=True)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"minv_chol_blocked" -s 2 -c 1 -f -o gpurun_out/prof_minv python /tmp/minv_one.py > gpurun_out/ncu_minv.log 2>&1
tail -2 gpurun_out/ncu_minv.log
