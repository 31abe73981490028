"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Mirrors unittest/parallel-rnea.cpp:21-55 and unittest/parallel-aba.cpp:21-55 (batched result ==
per-column serial result) with the north_star tolerance 1e-10 relative / 1e-12 absolute for FP64,
and extends it to crba / computeRNEADerivatives / computeABADerivatives.
"""
import numpy as np
import pytest

from conftest import MODEL_NAMES, assert_close, load_model, make_extra_models, random_inputs, structural_mask

pytestmark = pytest.mark.gpu

ALL_MODELS = MODEL_NAMES + ["mixed", "double_ff", "unaligned", "humanoid_hands", "wheeled"]


@pytest.fixture(scope="module")
def ctx(oracle_cls):
    import pinocchio_b200 as pb
    extra = make_extra_models()
    cache = {}

    def get(name):
        if name not in cache:
            model = extra[name] if name in extra else load_model(name)
            cache[name] = (model, pb.ModelPool(model), oracle_cls(model))
        return cache[name]
    yield get
    for _, pool, _ in cache.values():
        pool.close()


@pytest.mark.parametrize("name", ALL_MODELS)
@pytest.mark.parametrize("B", [1, 33, 128])
def test_rnea(ctx, name, B):
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, v, a = random_inputs(model, B, 11)
    tau = pb.rneaInParallel(1, pool, q, v, a)
    assert_close(tau, orc.rnea(q, v, a), what=f"rnea {name} B={B}")


@pytest.mark.parametrize("name", ALL_MODELS)
@pytest.mark.parametrize("B", [1, 33, 128])
def test_aba(ctx, name, B):
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, v, tau = random_inputs(model, B, 21)
    a = pb.abaInParallel(1, pool, q, v, tau)
    ref = orc.aba(q, v, tau)
    # ABA solves with M(q): judge the error at the scale of the solution column (Eigen isApprox semantics,
    # unittest/aba.cpp:154) in addition to the element-wise bound
    scale = np.abs(ref).max(axis=0, keepdims=True)
    assert_close(a, ref, rtol=1e-10, atol=1e-12 + 1e-10 * scale, what=f"aba {name} B={B}")


@pytest.mark.parametrize("name", ALL_MODELS)
@pytest.mark.parametrize("B", [1, 33, 128])
def test_crba(ctx, name, B):
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, _, _ = random_inputs(model, B, 31)
    M = pb.crbaInParallel(1, pool, q)
    ref = orc.crba(q, world=True)
    scale = np.abs(ref).max(axis=0, keepdims=True)
    assert_close(M, ref, rtol=1e-10, atol=1e-12 + 1e-12 * scale, what=f"crba {name} B={B}")
    # entries outside the tree sparsity (incl. the strictly-lower triangle) are exactly zero, as in a fresh Data
    assert not M[~structural_mask(model)].any()


@pytest.mark.parametrize("name", ALL_MODELS)
@pytest.mark.parametrize("B", [1, 33, 96])
def test_rnea_derivatives(ctx, name, B):
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, v, a = random_inputs(model, B, 41)
    dq, dv, da, tau = pb.computeRNEADerivativesInParallel(1, pool, q, v, a)
    rdq, rdv, rda, rtau = orc.rnea_derivatives(q, v, a)
    for got, ref, nm in ((dq, rdq, "dtau_dq"), (dv, rdv, "dtau_dv"), (da, rda, "dtau_da"), (tau, rtau, "tau")):
        scale = np.abs(ref).max(axis=0, keepdims=True)
        assert_close(got, ref, rtol=1e-10, atol=1e-12 + 1e-11 * scale, what=f"{nm} {name} B={B}")
    assert not da[~structural_mask(model)].any()
    assert not dq[~structural_mask(model, lower=True)].any() and not dv[~structural_mask(model, lower=True)].any()


@pytest.mark.parametrize("name", ALL_MODELS)
@pytest.mark.parametrize("B", [1, 33, 96])
def test_aba_derivatives(ctx, name, B):
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, v, tau = random_inputs(model, B, 51)
    dq, dv, dtau, ddq = pb.computeABADerivativesInParallel(1, pool, q, v, tau)
    rdq, rdv, rdtau, rddq = orc.aba_derivatives(q, v, tau)
    for got, ref, nm in ((dq, rdq, "ddq_dq"), (dv, rdv, "ddq_dv"), (dtau, rdtau, "ddq_dtau"), (ddq, rddq, "ddq")):
        scale = np.abs(ref).max(axis=0, keepdims=True)
        assert_close(got, ref, rtol=1e-10, atol=1e-12 + 1e-10 * scale, what=f"{nm} {name} B={B}")


def test_leading_dimension_and_batch_independence(ctx):
    """Column i of a batch equals the 1-column call (unittest/parallel-rnea.cpp:54), with ld > rows."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx("humanoid_random")
    B = 70
    q, v, a = random_inputs(model, B, 61)
    big = np.zeros((model.nq + 5, B), order="F")
    big[:model.nq] = q
    qv = big[:model.nq]  # ld = nq + 5
    tau = pb.rneaInParallel(1, pool, qv, v, a)
    tau0 = pb.rneaInParallel(1, pool, q, v, a)
    assert np.array_equal(tau, tau0)
    for i in (0, 31, 32, 69):
        ti = pb.rneaInParallel(1, pool, q[:, i:i + 1], v[:, i:i + 1], a[:, i:i + 1])
        assert np.array_equal(ti[:, 0], tau[:, i])


def test_device_pointers_match_host(ctx):
    import torch
    import pinocchio_b200 as pb
    model, pool, orc = ctx("simple_humanoid_ff")
    B = 257
    q, v, a = random_inputs(model, B, 71)
    tq, tv, ta = (torch.from_numpy(np.ascontiguousarray(x.T)).cuda() for x in (q, v, a))
    tau_d = pb.rneaInParallel(1, pool, tq, tv, ta)
    tau_h = pb.rneaInParallel(1, pool, q, v, a)
    assert np.array_equal(tau_d.cpu().numpy().T, tau_h)
    M_d = pb.crbaInParallel(1, pool, tq)
    assert np.array_equal(M_d.cpu().numpy().T, pb.crbaInParallel(1, pool, q))


def test_fp32_mode_tolerance(ctx):
    """FP32 mode: tolerance measured against the FP64 oracle and stated here."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx("talos_reduced_ff")
    B = 64
    q, v, a = random_inputs(model, B, 81)
    tau32 = pb.rneaInParallel(1, pool, q.astype(np.float32), v.astype(np.float32), a.astype(np.float32))
    ref = orc.rnea(q, v, a)
    assert tau32.dtype == np.float32
    rel = np.abs(tau32 - ref).max() / np.abs(ref).max()
    assert rel < 5e-5, rel
    M32 = pb.crbaInParallel(1, pool, q.astype(np.float32))
    refM = orc.crba(q, world=True)
    assert np.abs(M32 - refM).max() / np.abs(refM).max() < 5e-5
    # nv % 4 == 0: the FP32 column blocks leave through TMA tensor stores (talos, nv 38, takes the LSU emitter above)
    model, pool, orc = ctx("humanoid_random")
    q, _, _ = random_inputs(model, 77, 83)
    M32 = pb.crbaInParallel(1, pool, q.astype(np.float32))
    refM = orc.crba(q, world=True)
    assert M32.dtype == np.float32 and np.abs(M32 - refM).max() / np.abs(refM).max() < 5e-5
    assert not M32[~structural_mask(model)].any(), "entries outside the tree sparsity must be exact zeros"


def test_error_behaviour(ctx):
    """Argument-size errors surface as ValueError (reference: std::invalid_argument, macros.hpp:185-223)."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx("manipulator")
    q, v, a = random_inputs(model, 4, 91)
    with pytest.raises(ValueError):
        pb.rneaInParallel(1, pool, q[:-1], v, a)
    with pytest.raises(ValueError):
        pb.rneaInParallel(1, pool, q, v[:, :3], a)
    with pytest.raises(ValueError):
        pb.abaInParallel(1, pool, q, v, a, a=np.zeros((model.nv + 1, 4), order="F"))
    # empty batch is a no-op
    out = pb.rneaInParallel(1, pool, q[:, :0], v[:, :0], a[:, :0])
    assert out.shape == (model.nv, 0)
    # the other entry points check sizes the same way
    with pytest.raises(ValueError):
        pb.nonLinearEffectsInParallel(1, pool, q, v[:-1])
    with pytest.raises(ValueError):
        pb.computeMinverseInParallel(1, pool, q, Minv=np.zeros((model.nv * model.nv - 1, 4), order="F"))
    with pytest.raises(ValueError):
        pb.integrateInParallel(1, pool, q, v, qout=np.zeros((model.nq, 3), order="F"))
    with pytest.raises(ValueError):
        pb.abaEulerStepInParallel(1, pool, q, v[:, :2], a, 1e-3)
    assert pb.computeGeneralizedGravityInParallel(1, pool, q[:, :0]).shape == (model.nv, 0)


# ---- the callers' other needs on the same sweeps ---------------------------------------------------
@pytest.mark.parametrize("name", ALL_MODELS)
def test_nle_and_gravity(ctx, name):
    """nonLinearEffects == rnea(q, v, 0) (unittest/rnea.cpp:201-206), computeGeneralizedGravity == rnea(q, 0, 0)
    (unittest/rnea.cpp:225-228)."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, v, _ = random_inputs(model, 65, 31)
    assert_close(pb.nonLinearEffectsInParallel(1, pool, q, v), orc.nle(q, v), what=f"nle {name}")
    assert_close(pb.computeGeneralizedGravityInParallel(1, pool, q), orc.gravity(q), what=f"gravity {name}")


@pytest.mark.parametrize("name", ALL_MODELS)
def test_minverse(ctx, name):
    """computeMinverse: upper triangle of M^-1 (unittest/aba.cpp:265-301), zeros below; Minv == the ddq_dtau of
    computeABADerivatives (unittest/aba-derivatives.cpp:96-100)."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, v, tau = random_inputs(model, 37, 33)
    Minv = pb.computeMinverseInParallel(1, pool, q)
    ref = orc.minverse(q)
    assert_close(Minv, ref, atol=1e-12 + 1e-10 * np.abs(ref).max(), what=f"Minv {name}")
    nv = model.nv
    lower = ~np.triu(np.ones((nv, nv), dtype=bool)).reshape(-1, order="F")
    assert not Minv[lower].any()
    full = pb.computeABADerivativesInParallel(1, pool, q, v, tau)[2]
    assert_close(Minv[~lower], full[~lower], atol=1e-12 + 1e-10 * np.abs(ref).max(), what=f"Minv vs ddq_dtau {name}")


@pytest.mark.parametrize("name", ALL_MODELS)
@pytest.mark.parametrize("B", [1, 70])
def test_integrate(ctx, name, B):
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, v, _ = random_inputs(model, B, 35)
    assert_close(pb.integrateInParallel(1, pool, q, 0.3 * v), orc.integrate(q, 0.3 * v), what=f"integrate {name} B={B}")


@pytest.mark.parametrize("name", ALL_MODELS)
def test_aba_euler_step(ctx, name):
    """a = aba(q, v, tau); v += a dt; q = integrate(q, v dt) (examples/simulation-pendulum.py:153-157), three steps on the
    device against the same three steps with the oracle."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    q, v, tau = random_inputs(model, 40, 37)
    dt = 1e-3
    qg, vg, qo, vo = q.copy(order="F"), v.copy(order="F"), q.copy(order="F"), v.copy(order="F")
    for _ in range(3):
        qg, vg = pb.abaEulerStepInParallel(1, pool, qg, vg, tau, dt)
        a = orc.aba(qo, vo, tau)
        vo = np.asfortranarray(vo + dt * a)
        qo = orc.integrate(qo, dt * vo)
    s = max(1.0, np.abs(vo).max())
    assert_close(vg, vo, atol=1e-10 * s, what=f"euler v {name}")
    assert_close(qg, qo, atol=1e-11, what=f"euler q {name}")


def test_euler_step_device_resident_and_fp32(ctx):
    """Device pointers (torch CUDA tensors) and the FP32 mode of the new entry points."""
    import torch
    import pinocchio_b200 as pb
    model, pool, orc = ctx("talos_reduced_ff")
    q, v, tau = random_inputs(model, 96, 41)
    tq, tv, tt = (torch.from_numpy(np.ascontiguousarray(x.T)).cuda() for x in (q, v, tau))
    qn, vn = pb.abaEulerStepInParallel(1, pool, tq, tv, tt, 2e-3)
    a = orc.aba(q, v, tau)
    vo = v + 2e-3 * a
    qo = orc.integrate(q, 2e-3 * vo)
    assert_close(vn.cpu().numpy().T, vo, atol=1e-10 * max(1.0, np.abs(vo).max()), what="euler v (device)")
    assert_close(qn.cpu().numpy().T, qo, atol=1e-11, what="euler q (device)")
    q32 = pb.integrateInParallel(1, pool, q.astype(np.float32), (0.3 * v).astype(np.float32))
    assert np.abs(q32 - orc.integrate(q, 0.3 * v)).max() < 2e-6  # FP32 mode: ~10 ulp of float on O(1) coordinates
    g32 = pb.computeGeneralizedGravityInParallel(1, pool, q.astype(np.float32))
    gref = orc.gravity(q)
    assert np.abs(g32 - gref).max() < 5e-5 * np.abs(gref).max()


@pytest.mark.parametrize("name", ALL_MODELS)
@pytest.mark.parametrize("path", ["thread", "coop"])
def test_rnea_aba_both_paths(ctx, name, path, monkeypatch):
    """rneaInParallel / abaInParallel have two device paths — one configuration per thread (large batches) and G lanes per
    configuration (small batches, chosen by the batch size); BRBD_COOP_MAX_BATCH forces either. Both must match the oracle,
    and each other to rounding."""
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    monkeypatch.setenv("BRBD_COOP_MAX_BATCH", "0" if path == "thread" else "1000000")
    for B in (1, 33, 130):
        q, v, a = random_inputs(model, B, 51)
        tau = pb.rneaInParallel(1, pool, q, v, a)
        assert_close(tau, orc.rnea(q, v, a), atol=1e-12 * max(1.0, np.abs(tau).max()), what=f"rnea[{path}] {name} B={B}")
        ddq = pb.abaInParallel(1, pool, q, v, tau)
        ref = orc.aba(q, v, tau)
        assert_close(ddq, ref, atol=1e-12 + 1e-10 * np.abs(ref).max(), what=f"aba[{path}] {name} B={B}")


@pytest.mark.parametrize("name,pad", [("humanoid", 0), ("humanoid", 1), ("humanoid", 2), ("talos_reduced_ff", 6),
                                      ("simple_humanoid_ff", 0), ("simple_humanoid_ff", 1), ("humanoid_random", 0)])
@pytest.mark.parametrize("B", [1, 31, 77])
def test_crba_device_layouts(ctx, name, pad, B):
    """CRBA into device matrices with a padded leading dimension: even nv and even ld leave through TMA tensor stores
    (crba_tma_kernel), anything else through the LSU emitter (crba_tmem_kernel); the padding is never written and a partial
    last tile is clipped."""
    import torch
    import pinocchio_b200 as pb
    model, pool, orc = ctx(name)
    nn = model.nv * model.nv
    q, _, _ = random_inputs(model, B, 71)
    tq = torch.from_numpy(np.ascontiguousarray(q.T)).cuda()
    big = torch.full((B + 1, nn + pad), -7.0, dtype=torch.float64, device="cuda")
    pb.crbaInParallel(1, pool, tq, big[:B, :nn])
    torch.cuda.synchronize()
    got = big.cpu().numpy()
    assert_close(got[:B, :nn].T, orc.crba(q, world=True), atol=1e-11, what=f"crba {name} ld={nn + pad}")
    assert (got[:B, nn:] == -7.0).all() and (got[B] == -7.0).all(), "wrote outside the caller's block"
