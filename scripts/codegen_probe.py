#!/usr/bin/env python
"""Generate the specialised source of one algorithm for one model, check the host variant against the oracle (CPU),
cross-compile the device variant for sm_100a and report registers / spills / SASS opcode mix.
    python scripts/codegen_probe.py simple_humanoid_ff aba [--slots] [--nt 128] [--minb 1]"""
import argparse, ctypes, os, subprocess, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import load_model, make_extra_models, random_inputs
from pinocchio_b200 import _capi
from pinocchio_b200.codegen import codegen_source

ap = argparse.ArgumentParser()
ap.add_argument("model"); ap.add_argument("algo")
ap.add_argument("--slots", action="store_true"); ap.add_argument("--nt", type=int, default=128); ap.add_argument("--minb", type=int, default=1)
ap.add_argument("--no-device", action="store_true"); ap.add_argument("--direct", action="store_true")
args = ap.parse_args()
extra = make_extra_models()
model = extra[args.model] if args.model in extra else load_model(args.model)
out = os.path.join(ROOT, "gpurun_out", "codegen"); os.makedirs(out, exist_ok=True)
t0 = time.time()
src, info = codegen_source(model, args.algo, explicit_slots=args.slots, host=True)
print("trace+emit %.2fs" % (time.time() - t0), info)
hp = os.path.join(out, f"{args.model}_{args.algo}_host.cpp"); open(hp, "w").write(src)
so = hp[:-4] + ".so"
t0 = time.time()
subprocess.check_call(["/usr/bin/g++", "-O1", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, hp, "-lm"])
print("gcc %.1fs" % (time.time() - t0))
L = ctypes.CDLL(so)
fn = getattr(L, f"brbd_gen_{args.algo}_host")
from oracle import Oracle
orc = Oracle(model)
B = 16
q, v, x = random_inputs(model, B, 5)
nout = model.nv * model.nv if args.algo == "crba" else (3 * model.nv ** 2 + model.nv if "derivatives" in args.algo else model.nv)
res = np.zeros((nout, B), order="F")
rec = np.zeros(max(1, info["record_slots"])); park = np.zeros(max(1, info["park_slots"]))
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
for i in range(B):
    qi, vi, xi, oi = (np.ascontiguousarray(a[:, i]) for a in (q, v, x, res))
    rec[:] = np.nan; park[:] = np.nan
    fn(P(qi), P(vi), P(xi), P(oi), P(rec), P(park))
    res[:, i] = oi
if "derivatives" in args.algo:
    parts = orc.rnea_derivatives(q, v, x) if args.algo == "rnea_derivatives" else orc.aba_derivatives(q, v, x)
    ref = np.vstack(parts)
    if args.algo == "rnea_derivatives":  # dtau_da: upper triangle only in the reference; the engine's v1 code fills the same
        pass
else:
    ref = orc.aba(q, v, x) if args.algo == "aba" else (orc.crba(q, world=False) if args.algo == "crba" else orc.rnea(q, v, x))
print("max |err| / max |ref| vs oracle: %.2e" % (np.abs(res - ref).max() / np.abs(ref).max()))
if not args.no_device:
    src, info = codegen_source(model, args.algo, explicit_slots=args.slots, nt=args.nt, minb=args.minb, direct_io=args.direct)
    cu = os.path.join(out, f"{args.model}_{args.algo}.cu"); open(cu, "w").write(src)
    t0 = time.time()
    r = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-Xptxas", "-v", "-o", cu[:-3] + ".cubin", cu],
                       capture_output=True, text=True, timeout=300)
    print("nvcc %.1fs" % (time.time() - t0)); print(r.stderr[-900:])
    sass = subprocess.run(["cuobjdump", "-sass", cu[:-3] + ".cubin"], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in sass.splitlines():
        t = line.split()
        if len(t) > 1 and t[0].startswith("/*") and t[0].endswith("*/") and len(t[0]) in (8, 9, 10) and not t[1].startswith("/*"):
            op = t[2] if t[1].startswith("@") else t[1]
            ops[op.split(".")[0].rstrip(";")] += 1
    tot = sum(ops.values())
    print("SASS instructions", tot, "FP64", ops["DFMA"] + ops["DMUL"] + ops["DADD"], dict(ops.most_common(18)))
