"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Mirrors unittest/parallel-rnea.cpp:21-55 and unittest/parallel-aba.cpp:21-55 (batched result ==
per-column serial result) with the north_star tolerance 1e-10 relative / 1e-12 absolute for FP64,
and extends it to crba / computeRNEADerivatives / computeABADerivatives.
"""
import numpy as np
import pytest

from conftest import MODEL_NAMES, assert_close, load_model, make_extra_models, random_inputs

pytestmark = pytest.mark.gpu

ALL_MODELS = MODEL_NAMES + ["mixed", "double_ff"]


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


def structural_mask(model, lower=False):
    """(nv*nv,) bool, col-major: True where the tree sparsity allows a non-zero.
    Upper part: joint(row) is an ancestor-or-self of joint(col) (crba.hxx:94-95); with lower=True also the
    transposed entries (rnea-derivatives.hxx:433-438)."""
    nv = model.nv
    dof_joint = np.zeros(nv, dtype=int)
    for j in range(1, model.njoints):
        dof_joint[model.idx_vs[j]:model.idx_vs[j] + model.nvs[j]] = j
    anc = np.zeros((model.njoints, model.njoints), dtype=bool)  # anc[a, j]: a ancestor-or-self of j
    for j in range(1, model.njoints):
        a = j
        while a > 0:
            anc[a, j] = True
            a = model.parents[a]
    mask = anc[np.ix_(dof_joint, dof_joint)]
    if lower:
        mask = mask | mask.T
    return mask.reshape(-1, order="F")


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
