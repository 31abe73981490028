#!/usr/bin/env python
"""Device-resident throughput of all five algorithms on the config models, one JSON line per
(model, algorithm): configurations/s, algorithmic GB/s and FP64 TFLOP/s against the measured peaks.
    python scripts/bench_all.py [--models manipulator,simple_humanoid_ff] [--batch 65536] [--reps 5]
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
from oracle import Oracle, build_oracle

ap = argparse.ArgumentParser()
ap.add_argument("--models", default="manipulator,humanoid_random,simple_humanoid_ff,talos_reduced_ff")
ap.add_argument("--algos", default="rnea,aba,crba,rnea_derivatives,aba_derivatives,nle,gravity,minverse,integrate,euler_step")
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--generic", action="store_true", help="do not specialise the pool (generic kernels)")
args = ap.parse_args()
build_oracle()
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dt = torch.float64 if args.dtype == "f64" else torch.float32
es = 8 if args.dtype == "f64" else 4
fp64_peak = None
for name in args.models.split(","):
    model = load_model(name)
    pool = pb.ModelPool(model, [0])
    pool.set_stream(torch.cuda.current_stream().cuda_stream)  # CUDA events below are recorded on torch's stream
    if not args.generic:  # kernels generated for the model (the derivative programs only for small models)
        pool.specialize(["rnea", "aba", "crba", "rnea_derivatives", "aba_derivatives"], fp32=args.dtype != "f64")
    if fp64_peak is None:
        fp64_peak = pool.measure_fp64_peak()[0]
    orc = Oracle(model)
    nq, nv, B = model.nq, model.nv, args.batch
    q, v, a = random_inputs(model, B, 1)
    tq, tv, ta = (torch.from_numpy(np.ascontiguousarray(x.T)).to("cuda", dt) for x in (q, v, a))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    outs = {"vec": torch.empty((B, nv), dtype=dt, device="cuda"),
            "m1": torch.empty((B, nv * nv), dtype=dt, device="cuda"), "m2": torch.empty((B, nv * nv), dtype=dt, device="cuda"),
            "m3": torch.empty((B, nv * nv), dtype=dt, device="cuda"), "vec2": torch.empty((B, nv), dtype=dt, device="cuda"),
            "cfg": torch.empty((B, nq), dtype=dt, device="cuda")}
    calls = {
        "rnea": (lambda: pb.rneaInParallel(1, pool, tq, tv, ta, outs["vec"], async_=True), es * (nq + 3 * nv), "rnea"),
        "aba": (lambda: pb.abaInParallel(1, pool, tq, tv, ta, outs["vec"], async_=True), es * (nq + 3 * nv), "aba"),
        "crba": (lambda: pb.crbaInParallel(1, pool, tq, outs["m1"], async_=True), es * (nq + nv * nv), "crba_world"),
        "rnea_derivatives": (lambda: pb.computeRNEADerivativesInParallel(1, pool, tq, tv, ta, outs["m1"], outs["m2"], outs["m3"], async_=True),
                             es * (nq + 2 * nv + 3 * nv * nv), "rnea_derivatives"),
        "aba_derivatives": (lambda: pb.computeABADerivativesInParallel(1, pool, tq, tv, ta, outs["m1"], outs["m2"], outs["m3"], async_=True),
                            es * (nq + 2 * nv + 3 * nv * nv), "aba_derivatives"),
        # the callers' other needs; no operation count in the oracle -> HBM figures only
        "nle": (lambda: pb.nonLinearEffectsInParallel(1, pool, tq, tv, outs["vec"], async_=True), es * (nq + 2 * nv), None),
        "gravity": (lambda: pb.computeGeneralizedGravityInParallel(1, pool, tq, outs["vec"], async_=True), es * (nq + nv), None),
        "minverse": (lambda: pb.computeMinverseInParallel(1, pool, tq, outs["m1"], async_=True), es * (nq + nv * nv), None),
        "integrate": (lambda: pb.integrateInParallel(1, pool, tq, tv, outs["cfg"], async_=True), es * (2 * nq + nv), None),
        "euler_step": (lambda: pb.abaEulerStepInParallel(1, pool, tq, tv, ta, 1e-3, outs["cfg"], outs["vec2"], async_=True),
                       es * (2 * nq + 3 * nv), "aba"),
    }
    for algo in args.algos.split(","):
        fn, bytes_per, oname = calls[algo]
        flops = orc.count_flops(oname, q[:, 0], v[:, 0], a[:, 0])["flops"] if oname else 0
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            flush.zero_()  # L2 flush between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        print(json.dumps({"model": name, "nq": nq, "nv": nv, "algo": algo, "dtype": args.dtype, "batch": B, "ms": round(ms, 4),
                          "configs_per_s": B / (ms * 1e-3), "algorithmic_bytes_per_config": bytes_per, "algorithmic_flops_per_config": flops,
                          "GBs": bytes_per * B / (ms * 1e-3) / 1e9, "hbm_frac_of_measured": bytes_per * B / (ms * 1e-3) / 1e9 / hbm,
                          "fp64_TFLOPs": flops * B / (ms * 1e-3) / 1e12,
                          "fp64_frac_of_measured": flops * B / (ms * 1e-3) / fp64_peak}), flush=True)
    pool.close()
