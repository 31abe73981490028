"""GPU parity at the BASELINE.json batch sizes, through device pointers, against the OpenMP oracle.

* C2 (configs[1]): ABA + CRBA on 65 536 configurations of simple_humanoid + free-flyer — EVERY column is compared.
* C3 (configs[2]): computeRNEADerivatives + computeABADerivatives on 2^20 configurations of the 6-dof manipulator —
  first 2 048 + 2 048 random + the last columns.
* C4 (configs[3]): RNEA / ABA on 4 * 2^20 and CRBA on 2^20 configurations of talos (free-flyer + 32 revolute), sampled alike.
* every one-configuration-per-thread kernel over more than one round of its persistent grid (148 CTAs x <= 256 threads).
* every device path of every algorithm forced once (BRBD_*_V): the fallbacks must match the oracle too.
* a 58-dof model beyond the on-chip layouts of the tuned kernels.

Mirrors unittest/parallel-rnea.cpp:21-55 / parallel-aba.cpp:21-55 (batched == serial per column) at the sizes the
benchmark runs.
"""
import numpy as np
import pytest

from conftest import assert_close, load_model, make_extra_models, random_inputs, structural_mask

pytestmark = pytest.mark.gpu


def to_dev(*xs):
    import torch
    return tuple(torch.from_numpy(np.ascontiguousarray(x.T)).cuda() for x in xs)


def to_host(t):
    return np.ascontiguousarray(t.cpu().numpy()).T


@pytest.fixture(scope="module")
def ctx(oracle_cls):
    import pinocchio_b200 as pb
    extra = make_extra_models()
    cache = {}

    def get(name):
        if name not in cache:
            model = extra[name] if name in extra else load_model(name)
            cache[name] = (model, pb.ModelPool(model, [0]), oracle_cls(model))
        return cache[name]
    yield get
    for _, pool, _ in cache.values():
        pool.close()


def sample_columns(B, seed):
    rng = np.random.default_rng(seed)
    head = np.arange(min(2048, B))
    mid = rng.choice(B, size=min(2048, B), replace=False)
    return np.unique(np.concatenate([head, mid, np.arange(max(0, B - 33), B)]))


def test_bench_config_parity(ctx):
    """BASELINE configs[1] exactly as bench.py runs it: 65 536 x simple_humanoid + FF, ABA then CRBA, FP64, device
    pointers; all 65 536 columns against the oracle."""
    import torch
    import pinocchio_b200 as pb
    model, pool, orc = ctx("simple_humanoid_ff")
    B = 65536
    q, v, tau = random_inputs(model, B, 2024)
    tq, tv, tt = to_dev(q, v, tau)
    a = pb.abaInParallel(1, pool, tq, tv, tt)
    ref = orc.aba(q, v, tau)
    scale = np.abs(ref).max(axis=0, keepdims=True)
    assert_close(to_host(a), ref, rtol=1e-10, atol=1e-12 + 1e-10 * scale, what="aba C2 simple_humanoid_ff B=65536 (all columns)")
    M = pb.crbaInParallel(1, pool, tq)
    torch.cuda.synchronize()
    for c0 in range(0, B, 8192):  # the dense reference is 642 MB: compare in slabs
        refM = orc.crba(np.asfortranarray(q[:, c0:c0 + 8192]), world=True)
        got = to_host(M[c0:c0 + 8192])
        assert_close(got, refM, rtol=1e-10, atol=1e-12 + 1e-12 * np.abs(refM).max(axis=0, keepdims=True),
                     what=f"crba C2 simple_humanoid_ff B=65536 columns {c0}..")
        assert not got[~structural_mask(model)].any(), "entries outside the tree sparsity must be exact zeros"


def test_c3_manipulator_derivatives_1M(ctx):
    """BASELINE configs[2]: computeRNEADerivatives + computeABADerivatives, manipulator, 2^20 configurations."""
    import torch
    import pinocchio_b200 as pb
    model, pool, orc = ctx("manipulator")
    B = 1 << 20
    q, v, a = random_inputs(model, B, 31337)
    tq, tv, ta = to_dev(q, v, a)
    cols = sample_columns(B, 5)
    qs, vs, as_ = (np.asfortranarray(x[:, cols]) for x in (q, v, a))
    dq, dv, da, tau = pb.computeRNEADerivativesInParallel(1, pool, tq, tv, ta)
    rdq, rdv, rda, rtau = orc.rnea_derivatives(qs, vs, as_)
    tc = torch.from_numpy(cols).cuda()
    for got, ref, nm in ((dq, rdq, "dtau_dq"), (dv, rdv, "dtau_dv"), (da, rda, "dtau_da"), (tau, rtau, "tau")):
        scale = np.abs(ref).max(axis=0, keepdims=True)
        assert_close(to_host(got[tc]), ref, rtol=1e-10, atol=1e-12 + 1e-11 * scale, what=f"{nm} C3 manipulator B=2^20 (sampled)")
    del dq, dv, da
    torch.cuda.empty_cache()
    dq, dv, dtau, ddq = pb.computeABADerivativesInParallel(1, pool, tq, tv, ta)
    rdq, rdv, rdtau, rddq = orc.aba_derivatives(qs, vs, as_)
    for got, ref, nm in ((dq, rdq, "ddq_dq"), (dv, rdv, "ddq_dv"), (dtau, rdtau, "ddq_dtau"), (ddq, rddq, "ddq")):
        scale = np.abs(ref).max(axis=0, keepdims=True)
        assert_close(to_host(got[tc]), ref, rtol=1e-10, atol=1e-12 + 1e-10 * scale, what=f"{nm} C3 manipulator B=2^20 (sampled)")


def test_c4_talos_4M(ctx):
    """BASELINE configs[3] on one GPU: talos (free-flyer + 32 revolute), RNEA / ABA on 4 * 2^20 configurations, CRBA on 2^20."""
    import torch
    import pinocchio_b200 as pb
    model, pool, orc = ctx("talos_reduced_ff")
    B = 4 << 20
    q, v, a = random_inputs(model, B, 4242)
    tq, tv, ta = to_dev(q, v, a)
    cols = sample_columns(B, 6)
    tc = torch.from_numpy(cols).cuda()
    qs, vs, as_ = (np.asfortranarray(x[:, cols]) for x in (q, v, a))
    tau = pb.rneaInParallel(1, pool, tq, tv, ta)
    assert_close(to_host(tau[tc]), orc.rnea(qs, vs, as_), what="rnea C4 talos B=4*2^20 (sampled)")
    ddq = pb.abaInParallel(1, pool, tq, tv, ta)
    ref = orc.aba(qs, vs, as_)
    assert_close(to_host(ddq[tc]), ref, rtol=1e-10, atol=1e-12 + 1e-10 * np.abs(ref).max(axis=0, keepdims=True),
                 what="aba C4 talos B=4*2^20 (sampled)")
    del tau, ddq
    B2 = 1 << 20
    M = pb.crbaInParallel(1, pool, tq[:B2])
    cols2 = sample_columns(B2, 7)
    refM = orc.crba(np.asfortranarray(q[:, cols2]), world=True)
    assert_close(to_host(M[torch.from_numpy(cols2).cuda()]), refM, rtol=1e-10,
                 atol=1e-12 + 1e-12 * np.abs(refM).max(axis=0, keepdims=True), what="crba C4 talos B=2^20 (sampled)")


@pytest.mark.parametrize("name", ["manipulator", "humanoid_random", "mixed"])
def test_persistent_grid_rounds(ctx, name, monkeypatch):
    """More than two rounds of the persistent grid (148 CTAs x <= 256 threads) with a ragged tail, for every
    one-configuration-per-thread kernel; all columns compared."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    monkeypatch.setenv("BRBD_COOP_MAX_BATCH", "0")
    B = 148 * 256 * 2 + 1234 + 7
    q, v, a = random_inputs(model, B, 99)
    tq, tv, ta = to_dev(q, v, a)
    assert_close(to_host(pb.rneaInParallel(1, pool, tq, tv, ta)), orc.rnea(q, v, a), what=f"rnea rounds {name} B={B}")
    ref = orc.aba(q, v, a)
    assert_close(to_host(pb.abaInParallel(1, pool, tq, tv, ta)), ref, rtol=1e-10,
                 atol=1e-12 + 1e-10 * np.abs(ref).max(axis=0, keepdims=True), what=f"aba rounds {name} B={B}")
    refM = orc.crba(q, world=True)
    assert_close(to_host(pb.crbaInParallel(1, pool, tq)), refM, rtol=1e-10,
                 atol=1e-12 + 1e-12 * np.abs(refM).max(axis=0, keepdims=True), what=f"crba rounds {name} B={B}")
    qn = to_host(pb.integrateInParallel(1, pool, tq, 0.25 * tv))
    assert_close(qn, orc.integrate(q, 0.25 * v), what=f"integrate rounds {name} B={B}")


FORCED = [("BRBD_ABA_V", "v3", "aba"), ("BRBD_ABA_V", "dfs", "aba"), ("BRBD_ABA_V", "v1", "aba"),
          ("BRBD_CRBA_V", "tmem", "crba"), ("BRBD_CRBA_V", "dfs", "crba"), ("BRBD_CRBA_V", "v1", "crba"),
          ("BRBD_RNEA_V", "v1", "rnea"), ("BRBD_DRNEA_V", "v1", "drnea"), ("BRBD_DABA_V", "v1", "daba"),
          ("BRBD_MINV_V", "coop", "minv"), ("BRBD_MINV_V", "chol", "minv")]


@pytest.mark.parametrize("var,val,algo", FORCED)
@pytest.mark.parametrize("name", ["simple_humanoid_ff", "mixed", "double_ff"])
def test_forced_paths(ctx, name, var, val, algo, monkeypatch):
    """Every fallback kernel the launch code can pick is forced once and must match the oracle like the default path."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    monkeypatch.setenv("BRBD_COOP_MAX_BATCH", "0")
    monkeypatch.setenv(var, val)
    B = 300
    q, v, a = random_inputs(model, B, 17)
    n0 = pool.launch_count()
    if algo == "rnea":
        assert_close(pb.rneaInParallel(1, pool, q, v, a), orc.rnea(q, v, a), what=f"rnea[{val}] {name}")
    elif algo == "aba":
        ref = orc.aba(q, v, a)
        assert_close(pb.abaInParallel(1, pool, q, v, a), ref, rtol=1e-10,
                     atol=1e-12 + 1e-10 * np.abs(ref).max(axis=0, keepdims=True), what=f"aba[{val}] {name}")
    elif algo == "crba":
        refM = orc.crba(q, world=True)
        M = pb.crbaInParallel(1, pool, q)
        assert_close(M, refM, rtol=1e-10, atol=1e-12 + 1e-12 * np.abs(refM).max(axis=0, keepdims=True), what=f"crba[{val}] {name}")
        assert not M[~structural_mask(model)].any()
    elif algo == "drnea":
        got = pb.computeRNEADerivativesInParallel(1, pool, q, v, a)
        for g, r, nm in zip(got, orc.rnea_derivatives(q, v, a), ("dtau_dq", "dtau_dv", "dtau_da", "tau")):
            assert_close(g, r, rtol=1e-10, atol=1e-12 + 1e-11 * np.abs(r).max(axis=0, keepdims=True), what=f"{nm}[{val}] {name}")
    elif algo == "minv":  # the articulated-body computeMinverse of the cooperative kernel (the default is crba + Cholesky)
        ref = orc.minverse(q)
        got = pb.computeMinverseInParallel(1, pool, q)
        assert_close(got, ref, rtol=1e-10, atol=1e-12 + 1e-10 * np.abs(ref).max(axis=0, keepdims=True), what=f"Minv[{val}] {name}")
    else:
        got = pb.computeABADerivativesInParallel(1, pool, q, v, a)
        for g, r, nm in zip(got, orc.aba_derivatives(q, v, a), ("ddq_dq", "ddq_dv", "ddq_dtau", "ddq")):
            assert_close(g, r, rtol=1e-10, atol=1e-12 + 1e-10 * np.abs(r).max(axis=0, keepdims=True), what=f"{nm}[{val}] {name}")
    assert pool.launch_count() > n0


def test_model_beyond_the_on_chip_layouts(ctx):
    """humanoid_hands: 54 joints, nv = 58, depth 14, three branching joints on one root path.  Every entry point must
    accept it (round 1 refused nv > 48 and returned EINVAL from the cooperative derivative kernels)."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx("humanoid_hands")
    for B, coop in ((1, "1000000"), (97, "0"), (97, "1000000")):
        import os
        os.environ["BRBD_COOP_MAX_BATCH"] = coop
        try:
            q, v, a = random_inputs(model, B, 23)
            tau = pb.rneaInParallel(1, pool, q, v, a)
            assert_close(tau, orc.rnea(q, v, a), atol=1e-12 * max(1.0, np.abs(tau).max()), what=f"rnea humanoid_hands B={B}")
            ref = orc.aba(q, v, tau)
            assert_close(pb.abaInParallel(1, pool, q, v, tau), ref, rtol=1e-10, atol=1e-12 + 1e-10 * np.abs(ref).max(),
                         what=f"aba humanoid_hands B={B}")
        finally:
            del os.environ["BRBD_COOP_MAX_BATCH"]
    q, v, a = random_inputs(model, 40, 29)
    refM = orc.crba(q, world=True)
    assert_close(pb.crbaInParallel(1, pool, q), refM, rtol=1e-10, atol=1e-12 + 1e-12 * np.abs(refM).max(), what="crba humanoid_hands")
    for g, r, nm in zip(pb.computeRNEADerivativesInParallel(1, pool, q, v, a), orc.rnea_derivatives(q, v, a),
                        ("dtau_dq", "dtau_dv", "dtau_da", "tau")):
        assert_close(g, r, rtol=1e-10, atol=1e-12 + 1e-11 * np.abs(r).max(), what=f"{nm} humanoid_hands")
    tau = orc.rnea(q, v, a)
    for g, r, nm in zip(pb.computeABADerivativesInParallel(1, pool, q, v, tau), orc.aba_derivatives(q, v, tau),
                        ("ddq_dq", "ddq_dv", "ddq_dtau", "ddq")):
        assert_close(g, r, rtol=1e-10, atol=1e-12 + 1e-10 * np.abs(r).max(), what=f"{nm} humanoid_hands")
    mref = orc.minverse(q)
    assert_close(pb.computeMinverseInParallel(1, pool, q), mref, atol=1e-12 + 1e-10 * np.abs(mref).max(), what="Minv humanoid_hands")
    assert_close(pb.integrateInParallel(1, pool, q, 0.3 * v), orc.integrate(q, 0.3 * v), what="integrate humanoid_hands")


def test_fp32_solves_tolerance(ctx):
    """FP32 mode of the solves: the stated tolerance, measured as max |err| / max |ref| against the FP64 oracle.
    ABA / computeABADerivatives lose three more digits than the products on talos (worst-conditioned joint-space inertia of the
    config models): 9e-5 / 3.3e-4 measured; asserted at 5e-4 / 2e-3.  humanoid_random: 2.1e-6 / 2.4e-6 -> 2e-5."""
    import pinocchio_b200 as pb
    for name, tol_aba, tol_d in (("talos_reduced_ff", 5e-4, 2e-3), ("humanoid_random", 2e-5, 2e-5)):
        model, pool, orc = ctx(name)
        q, v, tau = random_inputs(model, 256, 87)
        f = lambda x: np.asfortranarray(x.astype(np.float32))
        a32 = pb.abaInParallel(1, pool, f(q), f(v), f(tau))
        ref = orc.aba(q, v, tau)
        assert a32.dtype == np.float32
        rel = np.abs(a32 - ref).max() / np.abs(ref).max()
        assert rel < tol_aba, (name, rel)
        got = pb.computeABADerivativesInParallel(1, pool, f(q), f(v), f(tau))
        for g, r, nm in zip(got, orc.aba_derivatives(q, v, tau), ("ddq_dq", "ddq_dv", "ddq_dtau", "ddq")):
            rel = np.abs(g - r).max() / np.abs(r).max()
            assert rel < tol_d, (name, nm, rel)
        got = pb.computeRNEADerivativesInParallel(1, pool, f(q), f(v), f(tau))
        for g, r, nm in zip(got, orc.rnea_derivatives(q, v, tau), ("dtau_dq", "dtau_dv", "dtau_da", "tau")):
            rel = np.abs(g - r).max() / np.abs(r).max()
            assert rel < 5e-5, (name, nm, rel)


def test_torch_stream_ordering(ctx):
    """Device tensors produced on torch's current stream are consumed without an explicit synchronisation: the call must run
    on that stream (the pool's private non-blocking stream raced with the producer)."""
    import torch
    import pinocchio_b200 as pb
    model, pool, orc = ctx("humanoid_random")
    pool.set_stream(None)
    B = 20000
    q, v, a = random_inputs(model, B, 3)
    tq, tv, ta = to_dev(q, v, a)
    side = torch.cuda.Stream()
    for stream in (torch.cuda.current_stream(), side):
        with torch.cuda.stream(stream):
            # a long producer right before the call: scale the inputs on the device, no host sync
            big = torch.randn(64 << 20, device="cuda")
            for _ in range(4):
                big = big * 1.0001
            v2 = tv * 2.0
            tau = pb.rneaInParallel(1, pool, tq, v2, ta, async_=True)
            out = tau * 1.0  # consumer on the same stream
        stream.synchronize()
        assert_close(to_host(out), orc.rnea(q, 2.0 * v, a), what="rnea after an unsynchronised producer")
    with pytest.raises(ValueError):
        pb.rneaInParallel(1, pool, tq.cpu().numpy().T, tv, ta)  # host / device mix


def test_pool_surface(ctx):
    """ModelPoolTpl::getModel / getModels / getData / getDatas / resize / update (pool/model.hpp:60-131)."""
    import pinocchio_b200 as pb
    model, _, orc = ctx("manipulator")
    pool = pb.ModelPool(model, [0])
    assert pool.size() == 1 and pool.getModel(0) is model and pool.getModels() == [model]
    assert pool.getData(0)["device"] == 0 and len(pool.getDatas()) == 1
    with pytest.raises(ValueError):
        pool.getModel(1)
    flat = pool.flat_model()
    assert flat["njoints"] == model.njoints and np.array_equal(flat["parents"], np.asarray(model.parents))
    q, v, a = random_inputs(model, 50, 1)
    t0 = pb.rneaInParallel(1, pool, q, v, a)
    pool.resize([0])
    assert pool.size() == 1
    assert np.array_equal(pb.rneaInParallel(1, pool, q, v, a), t0)
    # update: heavier links change the result, the old model gives the old result back
    m2 = load_model("manipulator")
    for Y in m2.inertias[1:]:
        Y.mass *= 2.0
        Y.sym = Y.sym * 2.0
    pool.update(m2)
    t2 = pb.rneaInParallel(1, pool, q, v, a)
    from oracle import Oracle
    assert_close(t2, Oracle(m2).rnea(q, v, a), what="rnea after pool.update")
    assert not np.allclose(t2, t0)
    pool.update(model)
    assert np.array_equal(pb.rneaInParallel(1, pool, q, v, a), t0)
    with pytest.raises(ValueError):
        pb.ModelPool(model, [0, 0])
    pool.close()


def test_multi_device_pool(ctx, monkeypatch):
    """A pool over several devices shards a host-pointer call itself (contiguous column ranges, no exchange).  With the same
    kernel on every shard (the launch picks the small-batch cooperative kernels by the PER-DEVICE batch, so they are switched
    off here) the results are bit-identical to the single-device pool; with the default selection they agree to rounding.
    Needs >= 2 GPUs."""
    import torch
    import pinocchio_b200 as pb
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    model, pool1, orc = ctx("simple_humanoid_ff")
    n = min(4, torch.cuda.device_count())
    pooln = pb.ModelPool(model, list(range(n)))
    assert pooln.size() == n and pooln.devices() == list(range(n))
    q, v, a = random_inputs(model, 9000, 12)
    assert_close(pb.rneaInParallel(1, pooln, q, v, a), orc.rnea(q, v, a), what="rnea multi-device pool")
    monkeypatch.setenv("BRBD_COOP_MAX_BATCH", "0")
    for B in (1, 5, 9000, 70001):
        q, v, a = random_inputs(model, B, 13)
        assert np.array_equal(pb.rneaInParallel(1, pooln, q, v, a), pb.rneaInParallel(1, pool1, q, v, a))
        assert np.array_equal(pb.abaInParallel(1, pooln, q, v, a), pb.abaInParallel(1, pool1, q, v, a))
    q, v, a = random_inputs(model, 9000, 14)
    assert np.array_equal(pb.crbaInParallel(1, pooln, q), pb.crbaInParallel(1, pool1, q))
    with pytest.raises(ValueError):
        pb.rneaInParallel(1, pooln, *to_dev(q, v, a))  # device pointers need a single-device pool
    pooln.resize([1])
    assert pooln.devices() == [1]
    assert np.array_equal(pb.rneaInParallel(1, pooln, q, v, a), pb.rneaInParallel(1, pool1, q, v, a))
    pooln.close()


SPEC_MODELS = ["manipulator", "humanoid", "humanoid_random", "simple_humanoid_ff", "talos_reduced_ff", "mixed", "double_ff", "unaligned",
               "humanoid_hands", "wheeled"]


@pytest.mark.parametrize("name", SPEC_MODELS)
def test_specialized_kernels(ctx, name, monkeypatch):
    """pool.specialize(): kernels generated for the model (codegen: the tree unrolled, constants folded), compiled with NVRTC.
    Same parity bar as the generic kernels, every column compared, ragged batch over several rounds of the grid; the generic
    path must still be what small batches take."""
    import pinocchio_b200 as pb
    model, _, orc = ctx(name)
    pool = pb.ModelPool(model, [0])
    pool.specialize(["rnea", "aba", "crba"], min_batch=1)
    assert set(pool.specialized()) == {"rnea", "aba", "crba"}
    monkeypatch.setenv("BRBD_CRBA_V", "gen")  # the launch keeps the hand-written CRBA above 24 dofs: force the generated one
    for B in (1, 257, 148 * 256 + 4321):
        q, v, a = random_inputs(model, B, 77)
        tq, tv, ta = to_dev(q, v, a)
        n0 = pool.launch_count()
        tau = to_host(pb.rneaInParallel(1, pool, tq, tv, ta))
        assert_close(tau, orc.rnea(q, v, a), atol=1e-12 * max(1.0, np.abs(tau).max()), what=f"rnea[generated] {name} B={B}")
        ddq = to_host(pb.abaInParallel(1, pool, tq, tv, ta))
        ref = orc.aba(q, v, a)
        assert_close(ddq, ref, rtol=1e-10, atol=1e-12 + 1e-10 * np.abs(ref).max(axis=0, keepdims=True), what=f"aba[generated] {name} B={B}")
        assert pool.launch_count() == n0 + 2
        if B <= 257 or model.nv <= 40:
            import torch
            nn = model.nv * model.nv
            big = torch.full((B + 1, nn + 3), -7.0, dtype=torch.float64, device="cuda")  # padded leading dimension + canary
            pb.crbaInParallel(1, pool, tq, big[:B, :nn])
            torch.cuda.synchronize()
            got = big.cpu().numpy()
            refM = orc.crba(q, world=True)
            assert_close(got[:B, :nn].T, refM, rtol=1e-10, atol=1e-12 + 1e-12 * np.abs(refM).max(axis=0, keepdims=True),
                         what=f"crba[generated] {name} B={B}")
            assert not got[:B, :nn].T[~structural_mask(model)].any(), "entries outside the tree sparsity must be exact zeros"
            assert (got[:B, nn:] == -7.0).all() and (got[B] == -7.0).all(), "wrote outside the caller's block"
    # host pointers, nle / gravity / Euler step ride on the same kernels
    q, v, a = random_inputs(model, 300, 78)
    assert_close(pb.nonLinearEffectsInParallel(1, pool, q, v), orc.nle(q, v), atol=1e-12 * max(1.0, np.abs(orc.nle(q, v)).max()),
                 what=f"nle[generated] {name}")
    qn, vn = pb.abaEulerStepInParallel(1, pool, q, v, a, 1e-3)
    vo = v + 1e-3 * orc.aba(q, v, a)
    assert_close(vn, vo, atol=1e-10 * max(1.0, np.abs(vo).max()), what=f"euler[generated] {name}")
    # update() drops the specialised kernels (they were generated for the previous model)
    pool.update(model)
    assert pool.specialized() == []
    pool.close()


def test_specialized_fp32(ctx):
    import pinocchio_b200 as pb
    model, _, orc = ctx("humanoid_random")
    pool = pb.ModelPool(model, [0])
    pool.specialize(["rnea", "aba"], fp32=True, min_batch=1)
    q, v, a = random_inputs(model, 500, 79)
    f = lambda x: np.asfortranarray(x.astype(np.float32))
    t32 = pb.rneaInParallel(1, pool, f(q), f(v), f(a))
    ref = orc.rnea(q, v, a)
    assert t32.dtype == np.float32 and np.abs(t32 - ref).max() / np.abs(ref).max() < 5e-5
    a32 = pb.abaInParallel(1, pool, f(q), f(v), f(a))
    ref = orc.aba(q, v, a)
    assert np.abs(a32 - ref).max() / np.abs(ref).max() < 2e-5
    pool.close()


@pytest.mark.parametrize("name", ["manipulator", "mixed", "wheeled", "unaligned"])
def test_specialized_derivative_kernels(ctx, name):
    """computeRNEADerivatives / computeABADerivatives generated for a small model (the engine's generic per-thread algorithm
    traced on the recording scalar; BASELINE configs[2] runs the 6-dof manipulator through them): same parity bar, exact zeros
    outside the tree sparsity, optional tau / ddq output."""
    import pinocchio_b200 as pb
    model, _, orc = ctx(name)
    pool = pb.ModelPool(model, [0])
    pool.specialize(["rnea_derivatives", "aba_derivatives"], min_batch=1)
    assert set(pool.specialized()) == {"rnea_derivatives", "aba_derivatives"}
    for B in (1, 300, 148 * 256 * 2 + 77):
        q, v, a = random_inputs(model, B, 91)
        tq, tv, ta = to_dev(q, v, a)
        n0 = pool.launch_count()
        got = pb.computeRNEADerivativesInParallel(1, pool, tq, tv, ta)
        ref = orc.rnea_derivatives(q, v, a)
        for g, r, nm in zip(got, ref, ("dtau_dq", "dtau_dv", "dtau_da", "tau")):
            assert_close(to_host(g), r, rtol=1e-10, atol=1e-12 + 1e-11 * np.abs(r).max(axis=0, keepdims=True), what=f"{nm}[generated] {name} B={B}")
        assert not to_host(got[2])[~structural_mask(model)].any()
        assert not to_host(got[0])[~structural_mask(model, lower=True)].any()
        got = pb.computeABADerivativesInParallel(1, pool, tq, tv, ta)
        ref = orc.aba_derivatives(q, v, a)
        for g, r, nm in zip(got, ref, ("ddq_dq", "ddq_dv", "ddq_dtau", "ddq")):
            assert_close(to_host(g), r, rtol=1e-10, atol=1e-12 + 1e-10 * np.abs(r).max(axis=0, keepdims=True), what=f"{nm}[generated] {name} B={B}")
        assert pool.launch_count() == n0 + 2
    pool.close()


@pytest.mark.parametrize("name", ["simple_humanoid_ff", "talos_reduced_ff", "humanoid_random", "mixed", "humanoid_hands"])
@pytest.mark.parametrize("mode", ["bulk", "lsu", "compact"])
def test_generated_crba_store_modes(ctx, name, mode, monkeypatch):
    """The three ways the generated CRBA writes M (codegen.cu): one asynchronous bulk copy per lane and column group (the default
    above 24 dofs), coalesced stores by the warp, compact staging of the structural pattern.  Every caller layout: dense and
    padded leading dimensions of either parity, a base pointer 8 (4) bytes off a 16-byte boundary — the bulk copies' alignment
    shift — FP64 and FP32, ragged batches, a canary around the block."""
    import torch
    import pinocchio_b200 as pb
    model, _, orc = ctx(name)
    nv, nn = model.nv, model.nv * model.nv
    monkeypatch.setenv("BRBD_GEN_CRBA_MODE", mode)
    monkeypatch.setenv("BRBD_CRBA_V", "gen")
    mask = structural_mask(model)
    for fp32 in (False, True):
        pool = pb.ModelPool(model, [0])
        pool.specialize(["crba"], fp32=fp32, min_batch=1)
        dt = torch.float32 if fp32 else torch.float64
        for B, pad, shift in ((1, 0, 0), (33, 1, 1), (148 * 96 + 45, 0, 0), (1000, 2, 3), (517, 3, 1)):
            q, _, _ = random_inputs(model, B, 5 + B)
            refM = orc.crba(q, world=True)
            tq = torch.from_numpy(np.ascontiguousarray(q.T)).to(dt).cuda()
            flat = torch.full(((B + 1) * (nn + pad) + 8,), -7.0, dtype=dt, device="cuda")
            blk = flat[shift:shift + B * (nn + pad)].view(B, nn + pad)
            pb.crbaInParallel(1, pool, tq, blk[:, :nn])
            torch.cuda.synchronize()
            got = flat.cpu().numpy().astype(np.float64)
            G = got[shift:shift + B * (nn + pad)].reshape(B, nn + pad)
            what = f"crba[generated, {mode}, {'fp32' if fp32 else 'fp64'}] {name} B={B} pad={pad} shift={shift}"
            if fp32:
                assert np.abs(G[:, :nn].T - refM).max() <= 2e-5 * np.abs(refM).max(), what
            else:
                assert_close(G[:, :nn].T, refM, rtol=1e-10, atol=1e-12 + 1e-12 * np.abs(refM).max(axis=0, keepdims=True), what=what)
            assert not G[:, :nn].T[~mask].any(), what + ": entries outside the tree sparsity must be exact zeros"
            assert (G[:, nn:] == -7.0).all() and (got[:shift] == -7.0).all() and (got[shift + B * (nn + pad):] == -7.0).all(), what + ": wrote outside the block"
        pool.close()


@pytest.mark.parametrize("name", ["simple_humanoid_ff", "talos_reduced_ff", "manipulator", "mixed", "double_ff", "wheeled"])
def test_crba_packed(ctx, name):
    """brbd_crba_packed_batch: the entries of crba's result inside the structural pattern (brbd_model_crba_pattern), column-major;
    device and host pointers, padded leading dimension, FP32; expanding it gives the dense result of brbd_crba_batch exactly."""
    import torch
    import pinocchio_b200 as pb
    model, _, orc = ctx(name)
    nv = model.nv
    pool = pb.ModelPool(model, [0])
    rows, cols = pool.crbaPattern()
    nnz = len(rows)
    key = cols.astype(np.int64) * nv + rows
    assert np.array_equal(np.flatnonzero(structural_mask(model)), key)  # the pattern is the tree sparsity of crba.hxx:94-95, column-major
    for B, pad in ((1, 0), (300, 5), (148 * 480 + 99, 0)):
        q, _, _ = random_inputs(model, B, 17 + B)
        refM = orc.crba(q, world=True)
        assert not np.delete(refM, key, axis=0).any()  # and covers every non-zero of the reference
        tq, = to_dev(q)
        big = torch.full((B + 1, nnz + pad), -7.0, dtype=torch.float64, device="cuda")
        pb.crbaPackedInParallel(1, pool, tq, big[:B, :nnz])
        torch.cuda.synchronize()
        got = big.cpu().numpy()
        assert_close(got[:B, :nnz].T, refM[key], rtol=1e-10, atol=1e-12 + 1e-12 * np.abs(refM).max(axis=0, keepdims=True), what=f"crba packed {name} B={B}")
        assert (got[:B, nnz:] == -7.0).all() and (got[B] == -7.0).all(), "wrote outside the caller's block"
        if B <= 300:
            Ph = pb.crbaPackedInParallel(1, pool, q)  # host pointers
            assert np.array_equal(Ph, got[:B, :nnz].T)
            dense = pb.crbaInParallel(1, pool, q)
            assert_close(pb.expandPackedCrba(pool, Ph), dense, rtol=1e-12, atol=1e-13 * np.abs(dense).max(), what=f"expand(packed) vs dense {name}")
            P32 = pb.crbaPackedInParallel(1, pool, np.asfortranarray(q.astype(np.float32)))
            assert P32.dtype == np.float32 and np.abs(P32 - refM[key]).max() <= 2e-5 * np.abs(refM).max()
    pool.close()


@pytest.mark.parametrize("name", ["simple_humanoid_ff", "talos_reduced_ff", "manipulator", "humanoid_hands", "mixed"])
def test_minverse_large(ctx, name, monkeypatch):
    """computeMinverse at bench sizes: crba into the device's work buffer in chunks of 32 768 configurations, then the dense
    Cholesky inversion (minv_chol.cuh); sampled columns against the oracle's articulated-body Minv, upper triangle + exact zeros,
    padded leading dimension, the generated and the generic CRBA underneath, FP32."""
    import torch
    import pinocchio_b200 as pb
    model, _, orc = ctx(name)
    nv, nn = model.nv, model.nv * model.nv
    B = 70001
    q, _, _ = random_inputs(model, B, 31)
    cols = sample_columns(B, 5)
    ref = orc.minverse(np.asfortranarray(q[:, cols]))
    tq, = to_dev(q)
    low = np.tril(np.ones((nv, nv), dtype=bool), -1).reshape(-1, order="F")
    for spec in (False, True):
        pool = pb.ModelPool(model, [0])
        if spec:
            pool.specialize(["crba"])
        big = torch.full((B + 1, nn + 1), -7.0, dtype=torch.float64, device="cuda")
        pb.computeMinverseInParallel(1, pool, tq, big[:B, :nn])
        torch.cuda.synchronize()
        got = big[torch.from_numpy(cols).cuda()].cpu().numpy()
        assert_close(got[:, :nn].T, ref, rtol=1e-10, atol=1e-12 + 1e-10 * np.abs(ref).max(axis=0, keepdims=True), what=f"Minv[chol, specialised={spec}] {name} B={B}")
        assert not got[:, :nn].T[low].any(), "strictly-lower part must be exact zeros"
        assert (got[:, nn] == -7.0).all() and (big[B].cpu().numpy() == -7.0).all(), "wrote outside the caller's block"
        if not spec:
            m32 = pb.computeMinverseInParallel(1, pool, np.asfortranarray(q[:, :500].astype(np.float32)))
            r32 = orc.minverse(np.asfortranarray(q[:, :500]))
            assert m32.dtype == np.float32 and np.abs(m32 - r32).max() <= (2e-3 if name == "talos_reduced_ff" else 2e-4) * np.abs(r32).max()
        pool.close()


@pytest.mark.parametrize("name", ["simple_humanoid_ff", "talos_reduced_ff", "humanoid_hands"])
def test_crba_host_threads_packed_transfer(ctx, name):
    """crbaInParallel on host blocks with num_threads >= 2: only the entries inside the tree sparsity cross PCIe and the host threads
    rebuild the caller's dense matrices (brbd_pool_set_host_threads, host_expand.cpp).  Same result as the plain dense copy — against
    the oracle, exact zeros outside the sparsity, a padded leading dimension with a canary, several chunks with a ragged tail, FP32."""
    import pinocchio_b200 as pb
    model, _, orc = ctx(name)
    nv, nn = model.nv, model.nv * model.nv
    pool = pb.ModelPool(model, [0])
    mask = structural_mask(model)
    B = 40000 + 123
    q, _, _ = random_inputs(model, B, 41)
    cols = sample_columns(B, 9)
    ref = orc.crba(np.asfortranarray(q[:, cols]), world=True)
    big = np.full((nn + 3, B), -7.0, order="F")
    n0 = pool.launch_count()
    pb.crbaInParallel(8, pool, q, big[:nn])
    assert pool.launch_count() - n0 >= 2  # several chunks
    got = big[:nn][:, cols]
    assert_close(got, ref, rtol=1e-10, atol=1e-12 + 1e-12 * np.abs(ref).max(axis=0, keepdims=True), what=f"crba[host threads] {name} B={B}")
    assert not big[:nn][~mask].any(), "entries outside the tree sparsity must be exact zeros"
    assert (big[nn:] == -7.0).all(), "wrote outside the caller's block"
    dense = pb.crbaInParallel(1, pool, q)  # the plain path (dense copy)
    assert_close(big[:nn], dense, rtol=1e-12, atol=1e-13 * np.abs(dense).max(), what=f"crba host threads vs dense copy {name}")
    q32 = np.asfortranarray(q[:, :9000].astype(np.float32))
    m32 = pb.crbaInParallel(4, pool, q32)
    r32 = orc.crba(np.asfortranarray(q[:, :9000]), world=True)
    assert m32.dtype == np.float32 and np.abs(m32 - r32).max() <= 2e-5 * np.abs(r32).max() and not m32[~mask].any()
    pool.close()
