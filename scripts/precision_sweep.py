#!/usr/bin/env python
"""BASELINE.json configs[4]: FP32 vs FP64 throughput / accuracy sweep over the batch size for all five algorithms (one GPU;
the batch shards embarrassingly, bench.py --gpus N gives the multi-GPU figures).  One JSON line per (model, algo, dtype, batch):
device-resident configurations/s and, for FP32, the error against the FP64 oracle on 256 columns, as max |err| / max |ref|.
    python scripts/precision_sweep.py [--models talos_reduced_ff,humanoid_random] [--max-log2 22]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
from oracle import Oracle, build_oracle

ap = argparse.ArgumentParser()
ap.add_argument("--models", default="talos_reduced_ff,humanoid_random")
ap.add_argument("--max-log2", type=int, default=22)
ap.add_argument("--generic", action="store_true", help="do not specialise the pool (generic kernels)")
args = ap.parse_args()
build_oracle()
for name in args.models.split(","):
    model = load_model(name)
    pool = pb.ModelPool(model, [0])
    pool.set_stream(torch.cuda.current_stream().cuda_stream)
    if not args.generic:  # kernels generated for the model, both precisions
        pool.specialize(["rnea", "aba", "crba"])
        pool.specialize(["rnea", "aba", "crba"], fp32=True)
    orc = Oracle(model)
    nq, nv = model.nq, model.nv
    nn = nv * nv
    base = 1 << 16
    q0, v0, a0 = random_inputs(model, base, 3)
    refs = {"rnea": [orc.rnea(q0[:, :256], v0[:, :256], a0[:, :256])], "aba": [orc.aba(q0[:, :256], v0[:, :256], a0[:, :256])],
            "crba": [orc.crba(q0[:, :256], world=True)], "rnea_derivatives": list(orc.rnea_derivatives(q0[:, :256], v0[:, :256], a0[:, :256]))[:3],
            "aba_derivatives": list(orc.aba_derivatives(q0[:, :256], v0[:, :256], a0[:, :256]))[:3]}
    for log2b in range(10, args.max_log2 + 1, 2):
        B = 1 << log2b
        for dt, tdt in (("f64", torch.float64), ("f32", torch.float32)):
            reps = (B + base - 1) // base
            tq, tv, ta = (torch.from_numpy(np.ascontiguousarray(np.tile(x, (1, reps))[:, :B].T)).to("cuda", tdt) for x in (q0, v0, a0))
            for algo in ("rnea", "aba", "crba", "rnea_derivatives", "aba_derivatives"):
                es = 8 if dt == "f64" else 4
                rows = {"rnea": [nv], "aba": [nv], "crba": [nn], "rnea_derivatives": [nn, nn, nn], "aba_derivatives": [nn, nn, nn]}[algo]
                if sum(rows) * es * B > 60e9:
                    continue  # keep the outputs within one GPU's HBM with room for the inputs
                bufs = [torch.empty((B, r), dtype=tdt, device="cuda") for r in rows]
                call = {"rnea": lambda: pb.rneaInParallel(1, pool, tq, tv, ta, bufs[0], async_=True),
                        "aba": lambda: pb.abaInParallel(1, pool, tq, tv, ta, bufs[0], async_=True),
                        "crba": lambda: pb.crbaInParallel(1, pool, tq, bufs[0], async_=True),
                        "rnea_derivatives": lambda: pb.computeRNEADerivativesInParallel(1, pool, tq, tv, ta, *bufs, async_=True),
                        "aba_derivatives": lambda: pb.computeABADerivativesInParallel(1, pool, tq, tv, ta, *bufs, async_=True)}[algo]
                call(); torch.cuda.synchronize()
                ts = []
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); call(); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                err = max(float(np.abs(b[:256].double().cpu().numpy().T - r).max() / np.abs(r).max()) for b, r in zip(bufs, refs[algo]))
                print(json.dumps({"model": name, "algo": algo, "dtype": dt, "batch": B, "ms": round(ms, 4), "configs_per_s": B / (ms * 1e-3),
                                  "max_err_over_max_ref": err}), flush=True)
                del bufs
            del tq, tv, ta
            torch.cuda.empty_cache()
    pool.close()
