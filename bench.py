#!/usr/bin/env python
"""bench.py — headline benchmark of the batched-dynamics hot path (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input: every algorithm of the chosen
configuration once.  One eval = one configuration through one algorithm.

    --config C2 (default)  BASELINE.json configs[1]: ABA + CRBA, 65 536 x simple_humanoid.urdf + free-flyer, FP64
    --config C3            configs[2]: computeRNEADerivatives + computeABADerivatives, 2^20 x manipulator (6-dof)
    --config C4            configs[3]: RNEA + ABA + CRBA, 4 * 2^20 x talos (free-flyer + 32 revolute)
    --scaling weak|strong  weak (default for C2 / C3): the batch above is PER GPU; strong (default for C4): the batch
                           above is the whole job, sharded over the ranks in contiguous column ranges
                           (reference benchmark being mirrored: benchmark/timings-parallel.cpp:38-64,81-112)

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # CPU arm: the OpenMP restatement of
                                                              # rneaInParallel/abaInParallel (oracle/), same batch
Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM, CUDA events on the
launching stream, max over ranks); `e2e` = same metric through the public host-pointer API with pinned HOST buffers
(H2D + kernels + D2H inside the timed region).  No NCCL anywhere: the path has no exchange step; the barrier and the
max-over-ranks of the timings go through a gloo group.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

L2_BYTES = 126 * 1024 * 1024

CONFIGS = {
    "C2": {"model": "simple_humanoid_ff", "batch": 65536, "algos": ["aba", "crba"], "scaling": "weak",
           "desc": "simple_humanoid.urdf + free-flyer, nq=36 nv=35"},
    "C3": {"model": "manipulator", "batch": 1 << 20, "algos": ["rnea_derivatives", "aba_derivatives"], "scaling": "weak",
           "desc": "buildModels::manipulator, 6-dof"},
    "C4": {"model": "talos_reduced_ff", "batch": 4 << 20, "algos": ["rnea", "aba", "crba"], "scaling": "strong",
           "desc": "talos_reduced.urdf + free-flyer (free-flyer + 32 revolute), nq=39 nv=38"},
}
ORACLE_KEY = {"rnea": "rnea", "aba": "aba", "crba": "crba_world", "rnea_derivatives": "rnea_derivatives",
              "aba_derivatives": "aba_derivatives"}
KERNEL_NAME = {"rnea": "rnea_dfs_kernel<double>", "aba": "aba_rr_kernel<double>", "crba": "crba_tma_kernel<double>",
               "rnea_derivatives": "rnea_derivatives_coop_kernel<double>", "aba_derivatives": "aba_derivatives_coop_kernel<double>"}
# dram__bytes_read.sum + dram__bytes_write.sum per launch, RECORDED from the committed `ncu --set full` capture named in
# `source` (not measured in this run); only quoted for the exact configuration / batch of that capture, null otherwise.
NCU_TRAFFIC = {
    ("C2", 65536, "crba"): {"bytes": 608.6e6, "source": "profiles/r2_v3_step_ncu_full.csv (brbd_gen_crba_0, 96 threads, bulk copies: 21.6 MB read + 587.0 MB written)"},
    ("C2", 65536, "crba:generic"): {"bytes": 644.2e6, "source": "profiles/r2_v1_step_ncu_full.csv (crba_tma_kernel<double,224,1>: 47.3 MB read + 596.9 MB written)"},
    ("C2", 65536, "aba:generic"): {"bytes": 814.2e6, "source": "profiles/r1_v8_step_ncu_full.csv (aba_rr_kernel<double,224>: 388.1 MB read + 426.1 MB written)"},
    ("C3", 1 << 20, "rnea_derivatives"): {"bytes": 1565.4e6, "source": "profiles/r2_C3_ncu_full.csv (brbd_gen_rnea_derivatives_0: 151.2 MB read + 1414.2 MB written, of which the 1.0 KB of local-memory spill stores per configuration)"},
    ("C3", 1 << 20, "aba_derivatives"): {"bytes": 3089.9e6, "source": "profiles/r2_C3_ncu_full.csv (brbd_gen_aba_derivatives_0: 161.4 MB read + 2928.5 MB written, of which the 3.9 KB of local-memory spill stores per configuration)"},
    ("C4", 4 << 20, "rnea"): {"bytes": 5826.4e6, "source": "profiles/r2_C4_ncu_full.csv (brbd_gen_rnea_0: 3866.1 MB read + 1960.3 MB written)"},
    ("C4", 4 << 20, "aba"): {"bytes": 33485.0e6, "source": "profiles/r2_C4_ncu_full.csv (brbd_gen_aba_0: 18776.0 MB read + 14709.1 MB written: the pass-3 records, 2.1 KB per configuration each way)"},
    ("C4", 4 << 20, "crba"): {"bytes": 49730.1e6, "source": "profiles/r2_C4_ncu_full.csv (brbd_gen_crba_0: 1323.9 MB read + 48406.2 MB written)"},
    ("C2", 65536, "aba"): {"bytes": 389.9e6, "source": "profiles/r2_v3_step_ncu_full.csv (brbd_gen_aba_0, 448 threads, L2 policies on the records: 233.4 MB read + 156.5 MB written; "
                                                       "the generic aba_rr_kernel: 814.2 MB, profiles/r1_v8_step_ncu_full.csv)"},
}


def load_model(name):
    from pinocchio_b200.model import Model
    with open(os.path.join(ROOT, "tests", "golden", "models", name + ".json")) as fh:
        return Model.from_json(fh.read())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_numbers(model, orc, algos):
    """Per-configuration algorithmic bytes (exact) and FLOPs: the operation count of the reference algorithm as
    restated in the oracle, from a run with the counting scalar that does NOT count products / sums with the structural 0 / 1
    of the joints' motion subspaces (the reference's specialised joints and the kernels never execute those)."""
    from conftest import random_inputs
    nq, nv = model.nq, model.nv
    q, v, a = random_inputs(model, 1, 3)
    by = {"rnea": 8 * (nq + 3 * nv), "aba": 8 * (nq + 3 * nv), "crba": 8 * (nq + nv * nv),
          "rnea_derivatives": 8 * (nq + 3 * nv + 3 * nv * nv), "aba_derivatives": 8 * (nq + 3 * nv + 3 * nv * nv)}
    out = {}
    for k in algos:
        fl = orc.count_flops(ORACLE_KEY[k], q[:, 0], v[:, 0], a[:, 0])
        out[k] = {"bytes": by[k], "flops": fl["flops"], "sincos": fl["sincos"]}
    return out


def host_threads():
    """Every hardware thread this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to
    its workers, which would time the CPU arm on one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def column_range(B, world, rank):
    """Contiguous column range of `rank` (the split brbd's own multi-device pool uses): pinocchio_b200.sharding."""
    from pinocchio_b200.sharding import column_range as cr
    return cr(int(B), int(world), int(rank))


def oracle_outputs(orc, algos, n):
    """Caller-owned output blocks of the CPU arm, allocated (and touched) once, as a caller of the reference reuses its Data."""
    nv, nn = orc.nv, orc.nv * orc.nv
    out = {}
    for k in algos:
        if k in ("rnea", "aba"):
            out[k] = np.zeros((nv, n), order="F")
        elif k == "crba":
            out[k] = np.zeros((nn, n), order="F")
        else:
            out[k] = tuple(np.zeros((nn, n), order="F") for _ in range(3)) + (np.zeros((nv, n), order="F"),)
    return out


def oracle_step(orc, algos, q, v, x, nthreads, outs):
    for k in algos:
        if k == "rnea":
            orc.rnea(q, v, x, nthreads=nthreads, out=outs[k])
        elif k == "aba":
            orc.aba(q, v, x, nthreads=nthreads, out=outs[k])
        elif k == "crba":
            orc.crba(q, nthreads=nthreads, world=True, out=outs[k])
        elif k == "rnea_derivatives":
            orc.rnea_derivatives(q, v, x, nthreads=nthreads, out=outs[k])
        else:
            orc.aba_derivatives(q, v, x, nthreads=nthreads, out=outs[k])


def cpu_sample(orc, model, algos, B, nthreads, target_s=12.0):
    """The step's algorithms with the OpenMP oracle on a bounded sample of the workload; (evals/s, sample size, seconds)."""
    from conftest import random_inputs
    n = min(B, 2048)
    q, v, x = random_inputs(model, n, 77)
    outs = oracle_outputs(orc, algos, n)
    oracle_step(orc, algos, q, v, x, nthreads, outs)  # warm-up
    t0 = time.perf_counter()
    oracle_step(orc, algos, q, v, x, nthreads, outs)
    probe = time.perf_counter() - t0
    n2 = int(min(B, max(n, n * target_s / max(probe, 1e-6))))
    q, v, x = random_inputs(model, n2, 78)
    outs = oracle_outputs(orc, algos, n2)
    t0 = time.perf_counter()
    oracle_step(orc, algos, q, v, x, nthreads, outs)
    dt = time.perf_counter() - t0
    return len(algos) * n2 / dt, n2, dt


def workload_text(cfg_name, cfg, B_rank, world, scaling):
    total = B_rank * world if scaling == "weak" else cfg["batch"]
    return (f"{cfg_name}: {' + '.join(cfg['algos'])} on {total} configurations "
            f"({B_rank} per GPU x {world} GPU(s), {scaling} scaling) of {cfg['model']} ({cfg['desc']}), FP64; "
            f"one eval = one configuration through one algorithm")


def run_reference(args, cfg_name, cfg, scaling):
    """CPU arm: the restated reference path (oracle port) with all host threads on the SAME batch as the GPU arm's rank 0,
    after a warm-up of at least 3 s (benchmark/timings-parallel.cpp:103)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import Oracle, build_oracle
    build_oracle()
    model = load_model(cfg["model"])
    orc = Oracle(model)
    nthreads = host_threads()
    from conftest import random_inputs
    world = max(1, args.gpus)
    B = args.batch or cfg["batch"]
    n = B if scaling == "weak" else (column_range(B, world, 0)[1] - column_range(B, world, 0)[0])
    algos = cfg["algos"]
    q, v, x = random_inputs(model, n, 5)
    outs = oracle_outputs(orc, algos, n)
    t_w, nwarm = time.perf_counter(), 0
    while nwarm < args.warmup or time.perf_counter() - t_w < 3.0:
        oracle_step(orc, algos, q, v, x, nthreads, outs)
        nwarm += 1
    times = []
    for s in range(args.steps):
        t0 = time.perf_counter()
        oracle_step(orc, algos, q, v, x, nthreads, outs)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = len(algos) * n / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": f"batched dynamics evals/sec ({' + '.join(algos)})", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": nwarm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(cfg_name, cfg, n, 1, "weak") + f"; CPU arm: the whole {n}-configuration batch of one "
                                                                           f"GPU rank per step, {nthreads} OpenMP threads",
                   "batch_per_gpu": n},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": nthreads, "kind": "port",
                         "sample": f"{n} configurations per step (the full per-GPU batch) through {' + '.join(algos)}, OpenMP "
                                   f"schedule(static), the restated rneaInParallel/abaInParallel driver (oracle/), {nthreads} threads, "
                                   f"{nwarm} warm-up steps (>= 3 s)"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=0, help="override the configuration's batch (per GPU if weak, total if strong)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg")
    ap.add_argument("--generic", action="store_true", help="do not specialise the pool for the model (generic kernels only)")
    args = ap.parse_args()
    cfg_name, cfg = args.config, CONFIGS[args.config]
    scaling = args.scaling or cfg["scaling"]
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args, cfg_name, cfg, scaling)

    import torch
    import pinocchio_b200 as pb
    from conftest import random_inputs

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")  # barrier + max of three timings; the data path has no collective
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    model = load_model(cfg["model"])
    algos = cfg["algos"]
    nq, nv = model.nq, model.nv
    Btot = args.batch or cfg["batch"]
    if scaling == "weak":
        B, c0 = Btot, 0
    else:
        c0, c1 = column_range(Btot, world, rank)
        B = c1 - c0
    pool = pb.ModelPool(model, [local_rank])
    spec = [] if args.generic else [k for k in algos if k in ("rnea", "aba", "crba") or (k.endswith("derivatives") and model.nv <= 16)]
    if spec:
        try:
            pool.specialize(spec)  # kernels generated for this model (codegen + NVRTC), outside the timed region like the pool itself
        except Exception as e:  # no NVRTC on this box: the generic kernels remain in use, and the line says so
            print(f"bench.py: pool.specialize failed ({e}); generic kernels", file=sys.stderr)
            spec = []
        spec = [k for k in spec if k in pool.specialized()]
    stream = torch.cuda.current_stream()
    pool.set_stream(stream.cuda_stream)

    # Rotating input sets so that consecutive steps never find their inputs in L2 (126 MB); one set when a single one
    # is already larger than L2
    in_bytes = 8 * B * (nq + 2 * nv)
    nsets = max(2, int(np.ceil(2.5 * L2_BYTES / in_bytes))) if in_bytes < 2 * L2_BYTES else 1
    sets = []
    for s in range(nsets):
        q, v, x = random_inputs(model, B, 1000 * rank + 10 * s + (c0 % 9973))
        sets.append(tuple(torch.from_numpy(np.ascontiguousarray(t.T)).to(dev) for t in (q, v, x)))
    outs = {}
    for k in algos:
        if k in ("rnea", "aba"):
            outs[k] = (torch.empty((B, nv), dtype=torch.float64, device=dev),)
        elif k == "crba":
            outs[k] = (torch.empty((B, nv * nv), dtype=torch.float64, device=dev),)
        else:
            outs[k] = tuple(torch.empty((B, nv * nv), dtype=torch.float64, device=dev) for _ in range(3)) + \
                      (torch.empty((B, nv), dtype=torch.float64, device=dev),)

    def launch(k, q, v, x):
        if k == "rnea":
            pb.rneaInParallel(1, pool, q, v, x, outs[k][0], async_=True)
        elif k == "aba":
            pb.abaInParallel(1, pool, q, v, x, outs[k][0], async_=True)
        elif k == "crba":
            pb.crbaInParallel(1, pool, q, outs[k][0], async_=True)
        elif k == "rnea_derivatives":
            pb.computeRNEADerivativesInParallel(1, pool, q, v, x, *outs[k], async_=True)
        else:
            pb.computeABADerivativesInParallel(1, pool, q, v, x, *outs[k], async_=True)

    def step(i, ev=None):
        q, v, x = sets[i % nsets]
        for n, k in enumerate(algos):
            if ev is not None:
                ev[n].record(stream)
            launch(k, q, v, x)
        if ev is not None:
            ev[len(algos)].record(stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = pool.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(algos) + 1)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i, evs[i])
    e1.record(stream)
    barrier()
    launches = pool.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    algo_ms = [float(np.mean([e[n].elapsed_time(e[n + 1]) for e in evs])) for n in range(len(algos))]
    red = max_over_ranks([total_ms] + algo_ms)
    total_ms, algo_ms = red[0], red[1:]
    ms_per_step = total_ms / args.steps
    evals_per_step = len(algos) * (B * world if scaling == "weak" else Btot)
    value = evals_per_step / (ms_per_step * 1e-3)

    # ---- end to end through the public API with pinned HOST buffers (H2D + kernels + D2H timed) ----
    e2e = None
    if not args.no_e2e:
        pool.set_stream(None)
        out_rows = {"rnea": nv, "aba": nv, "crba": nv * nv, "rnea_derivatives": 3 * nv * nv + nv, "aba_derivatives": 3 * nv * nv + nv}
        per_cfg_out = 8 * sum(out_rows[k] for k in algos)
        Be = int(min(B, max(4096, (6 << 30) // per_cfg_out)))  # host blocks capped at ~6 GB of pinned memory
        e2e_steps = max(3, min(args.steps, 5))
        e2e_threads = max(1, host_threads() // max(1, world)) if "crba" in spec else 1
        hin = [torch.from_numpy(np.ascontiguousarray(t[:Be].cpu().numpy())).pin_memory() for t in sets[0]]
        hq, hv, hx = (t.numpy().T for t in hin)
        houts = {k: [torch.empty(tuple(o[:Be].shape), dtype=torch.float64).pin_memory() for o in outs[k]] for k in algos}
        hviews = {k: [t.numpy().T for t in houts[k]] for k in algos}

        def e2e_step():
            acc = 0.0
            for k in algos:
                o = hviews[k]
                if k == "rnea":
                    pb.rneaInParallel(1, pool, hq, hv, hx, o[0])
                elif k == "aba":
                    pb.abaInParallel(1, pool, hq, hv, hx, o[0])
                elif k == "crba":
                    # num_threads as the reference arm uses them: all host cores (here they rebuild the dense matrices from
                    # the packed PCIe transfer; the reference spends them on the algorithm)
                    pb.crbaInParallel(e2e_threads, pool, hq, o[0])
                elif k == "rnea_derivatives":
                    pb.computeRNEADerivativesInParallel(1, pool, hq, hv, hx, *o)
                else:
                    pb.computeABADerivativesInParallel(1, pool, hq, hv, hx, *o)
                acc += float(o[0][0, 0])
            return acc

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks([time.perf_counter() - t0])[0]
        in_rows = {"rnea": nq + 2 * nv, "aba": nq + 2 * nv, "crba": nq, "rnea_derivatives": nq + 2 * nv, "aba_derivatives": nq + 2 * nv}
        e2e = {"value": len(algos) * Be * world * e2e_steps / e2e_s, "unit": "evals/s",
               "h2d_bytes_per_step": int(8 * Be * sum(in_rows[k] for k in algos)),
               "d2h_bytes_per_step": int(Be * per_cfg_out), "steps": e2e_steps, "batch_per_gpu": Be,
               "result_bytes_per_step": int(Be * per_cfg_out),  # what lands in the caller's blocks
               "host_threads": e2e_threads,
               "note": "pinned host buffers -> brbd_*_batch(BRBD_PTR_HOST): H2D + kernels + D2H, wall clock, max over ranks"
                       + (f"; crba with num_threads = {e2e_threads}: only the entries inside the tree sparsity cross PCIe and the host threads rebuild the dense matrices" if e2e_threads >= 2 and "crba" in algos else "")
                       + ("" if Be == B else f"; host blocks capped at {Be} configurations per GPU (6 GB of pinned memory)")}

        if e2e_threads >= 2 and "crba" in algos and Be >= 4096:
            try:  # bytes that actually cross PCIe: the pattern entries of M, not the dense matrices
                e2e["d2h_bytes_per_step"] = int(Be * (per_cfg_out - 8 * (nv * nv - len(pool.crbaPattern()[0]))))
            except Exception:
                pass
        # the same step with the opt-in packed CRBA output (brbd_crba_packed_batch: only the entries inside the structural
        # pattern travel; a side figure — the headline `e2e` above is the reference's dense layout)
        if "crba" in algos and "crba" in spec:
            try:
                nnz = len(pool.crbaPattern()[0])
                hp = torch.empty((Be, nnz), dtype=torch.float64).pin_memory()
                hpv = hp.numpy().T

                def packed_step():
                    acc = 0.0
                    for k in algos:
                        o = hviews[k]
                        if k == "rnea":
                            pb.rneaInParallel(1, pool, hq, hv, hx, o[0])
                        elif k == "aba":
                            pb.abaInParallel(1, pool, hq, hv, hx, o[0])
                        else:
                            pb.crbaPackedInParallel(1, pool, hq, hpv)
                            o = [hpv]
                        acc += float(o[0][0, 0])
                    return acc

                packed_step()
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    packed_step()
                barrier()
                ps = max_over_ranks([time.perf_counter() - t0])[0]
                e2e["packed_crba"] = {"value": len(algos) * Be * world * e2e_steps / ps, "unit": "evals/s",
                                      "d2h_bytes_per_step": int(Be * (per_cfg_out - 8 * (nv * nv - nnz))),
                                      "note": f"same step, CRBA through brbd_crba_packed_batch ({nnz} of {nv * nv} entries per configuration)"}
            except Exception as e:  # a side figure must not cost the line
                print(f"bench.py: packed CRBA leg failed ({e})", file=sys.stderr)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- rooflines of the step's kernels + CPU baseline (rank 0) ----------------------------------
    from oracle import Oracle, build_oracle
    build_oracle()
    orc = Oracle(model)
    alg = algorithmic_numbers(model, orc, algos)
    hbm_peak, peak_src = measured_peaks()
    fp64_peak, _ = pool.measure_fp64_peak()
    kern = {}
    step_sum = sum(algo_ms)
    for name, ms in zip(algos, algo_ms):
        gbs = alg[name]["bytes"] * B / (ms * 1e-3) / 1e9
        tfl = alg[name]["flops"] * B / (ms * 1e-3) / 1e12
        hbm_frac, fp64_frac = gbs / hbm_peak, tfl / (fp64_peak / 1e12)
        kern[name] = {"kernel": (f"brbd_gen_{name} (generated for the model, NVRTC)" if name in spec else KERNEL_NAME[name]), "ms_per_launch": ms, "share_of_step": ms / step_sum, "configs_per_s": B / (ms * 1e-3),
                      "algorithmic_bytes_per_config": alg[name]["bytes"], "algorithmic_flops_per_config": alg[name]["flops"],
                      "sincos_per_config": alg[name]["sincos"], "achieved_GBs": gbs, "hbm_frac": hbm_frac,
                      "achieved_fp64_TFLOPs": tfl, "fp64_frac_of_measured_dfma_peak": fp64_frac,
                      # the roofline that bounds the kernel is the slower of the two ceilings, i.e. the larger fraction
                      "bound": "fp64" if fp64_frac >= hbm_frac else "hbm"}

    def roofline_of(name):
        k = kern[name]
        tr = NCU_TRAFFIC.get((cfg_name, B, name if name in spec else name + ":generic"))
        if k["bound"] == "fp64":
            r = {"bound": "fp64", "kernel": k["kernel"], "achieved": k["achieved_fp64_TFLOPs"], "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                 "frac": k["fp64_frac_of_measured_dfma_peak"], "peak_source": "measured in this run (register-resident DFMA loop)"}
        else:
            r = {"bound": "hbm", "kernel": k["kernel"], "achieved": k["achieved_GBs"], "peak": hbm_peak, "unit": "GB/s", "frac": k["hbm_frac"],
                 "peak_source": peak_src}
        r["traffic"] = tr["bytes"] if tr else None
        r["traffic_source"] = tr["source"] if tr else None
        r["share_of_step"] = k["share_of_step"]
        return r

    order = sorted(algos, key=lambda n: -kern[n]["share_of_step"])
    line = {
        "metric": f"batched dynamics evals/sec ({' + '.join(algos)})", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(cfg_name, cfg, B, world, scaling), "batch_per_gpu": B,
                   "l2_policy": (f"inputs rotate over {nsets} resident sets ({nsets * in_bytes >> 20} MiB > 126 MiB L2)" if nsets > 1
                                 else f"one input set of {in_bytes >> 20} MiB (> 126 MiB L2)") + "; outputs stream through L2 every step",
                   "parallelism": f"batch sharded over {world} GPU(s) in contiguous column ranges, no collective (gloo barrier only)"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": e2e,
        "roofline": roofline_of(order[0]),          # the kernel with the largest share of the step
        "other_rooflines": [roofline_of(n) for n in order[1:]],
        "kernels": kern,
        "flop_count": "oracle counting scalar, structural 0 / 1 of the joints' motion subspaces not counted (oracle/rbd_oracle.hpp Counted)",
        "fp64_peak_measured_TFLOPs": fp64_peak / 1e12,
    }
    if world == 1 and not args.no_cpu:
        nthreads = host_threads()
        cpu_v, n2, dt = cpu_sample(orc, model, algos, B, nthreads)
        line["cpu_baseline"] = {"value": cpu_v, "unit": "evals/s", "cores": nthreads, "kind": "port",
                                "sample": f"{n2} configurations through {' + '.join(algos)} in {dt:.1f} s with the OpenMP "
                                          f"restatement of rneaInParallel/abaInParallel (oracle/), schedule(static)"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
