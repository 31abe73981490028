#!/usr/bin/env python
"""North-star batch sizes (>= 1M configurations per call, BASELINE.json configs[2..3]): device-resident throughput and parity
on a column sample.  Inputs are generated on the host in chunks with the reference's distributions (seeded), uploaded once;
a prefix and a random sample of the output columns are checked against the oracle at the parity tolerance.
    python scripts/large_batch_check.py [--quick]
One JSON line per (model, algorithm, batch)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
from oracle import Oracle, build_oracle

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
build_oracle()
CASES = [("manipulator", 1 << 20, ["rnea_derivatives", "aba_derivatives", "rnea", "aba", "crba"]),
         ("talos_reduced_ff", 1 << 22, ["rnea", "aba"]),
         ("talos_reduced_ff", 1 << 20, ["crba", "rnea_derivatives"]),
         ("simple_humanoid_ff", 1 << 20, ["crba", "aba_derivatives", "euler_step"])]
if args.quick:
    CASES = [(m, b >> 4, a) for m, b, a in CASES]


def tiled_inputs(model, B, seed):
    """B columns: a 65536-column seeded block with the reference's distributions, tiled with a per-tile sign flip / scale so
    that columns differ (the kernels do not care; parity is checked on sampled columns against the oracle)."""
    base = 1 << 16
    q0, v0, a0 = random_inputs(model, min(B, base), seed)
    reps = (B + base - 1) // base
    q = np.tile(q0, (1, reps))[:, :B]
    scale = np.repeat(1.0 - 0.05 * (np.arange(reps) % 7), base)[:B]
    v = np.tile(v0, (1, reps))[:, :B] * scale
    a = np.tile(a0, (1, reps))[:, :B] * scale[::-1]
    return np.asfortranarray(q), np.asfortranarray(v), np.asfortranarray(a)


for name, B, algos in CASES:
    model = load_model(name)
    pool = pb.ModelPool(model, [0])
    pool.set_stream(torch.cuda.current_stream().cuda_stream)
    orc = Oracle(model)
    nq, nv = model.nq, model.nv
    q, v, a = tiled_inputs(model, B, 5)
    tq, tv, ta = (torch.from_numpy(np.ascontiguousarray(x.T)).cuda() for x in (q, v, a))
    rng = np.random.default_rng(9)
    cols = np.unique(np.concatenate([np.arange(min(B, 2048)), rng.integers(0, B, 2048), [B - 1]]))
    for algo in algos:
        nn = nv * nv
        outs = {"rnea": [nv], "aba": [nv], "crba": [nn], "rnea_derivatives": [nn, nn, nn, nv], "aba_derivatives": [nn, nn, nn, nv],
                "euler_step": [nq, nv]}[algo]
        bufs = [torch.empty((B, r), dtype=torch.float64, device="cuda") for r in outs]
        call = {"rnea": lambda: pb.rneaInParallel(1, pool, tq, tv, ta, bufs[0], async_=True),
                "aba": lambda: pb.abaInParallel(1, pool, tq, tv, ta, bufs[0], async_=True),
                "crba": lambda: pb.crbaInParallel(1, pool, tq, bufs[0], async_=True),
                "rnea_derivatives": lambda: pb.computeRNEADerivativesInParallel(1, pool, tq, tv, ta, *bufs, async_=True),
                "aba_derivatives": lambda: pb.computeABADerivativesInParallel(1, pool, tq, tv, ta, *bufs, async_=True),
                "euler_step": lambda: pb.abaEulerStepInParallel(1, pool, tq, tv, ta, 1e-3, bufs[0], bufs[1], async_=True)}[algo]
        call(); torch.cuda.synchronize()
        times = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        qs, vs, as_ = q[:, cols], v[:, cols], a[:, cols]
        ref = {"rnea": lambda: [orc.rnea(qs, vs, as_, nthreads=8)], "aba": lambda: [orc.aba(qs, vs, as_, nthreads=8)],
               "crba": lambda: [orc.crba(qs, nthreads=8, world=True)],
               "rnea_derivatives": lambda: list(orc.rnea_derivatives(qs, vs, as_, nthreads=8)),
               "aba_derivatives": lambda: list(orc.aba_derivatives(qs, vs, as_, nthreads=8)),
               "euler_step": lambda: (lambda vo: [orc.integrate(qs, 1e-3 * vo), vo])(vs + 1e-3 * orc.aba(qs, vs, as_, nthreads=8))}[algo]()
        worst = 0.0
        idx = torch.from_numpy(cols).cuda()
        for buf, r in zip(bufs, ref):
            got = buf.index_select(0, idx).cpu().numpy().T
            assert np.isfinite(got).all(), (name, algo)
            err = np.abs(got - r) / (1e-12 + 1e-10 * np.maximum(np.abs(r), np.abs(r).max() * (algo.endswith("derivatives") or algo == "aba" or algo == "euler_step")))
            worst = max(worst, float(err.max()))
        print(json.dumps({"model": name, "algo": algo, "batch": B, "ms": round(ms, 3), "configs_per_s": B / (ms * 1e-3),
                          "checked_columns": int(len(cols)), "worst_error_over_tolerance": round(worst, 4), "parity": bool(worst <= 1.0)}), flush=True)
        del bufs
        torch.cuda.empty_cache()
    pool.close()
    del tq, tv, ta
    torch.cuda.empty_cache()
