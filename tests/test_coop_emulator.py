"""CPU lane emulation of the warp-cooperative kernels (tests/cpp/coop_emu.cu): the per-configuration device code of
computeRNEADerivatives / computeABADerivatives (deriv_coop.cuh, aba_deriv_coop.cuh) compiled for the host, one thread
per lane, __syncwarp() = a barrier, against the oracle.  This checks the lane / phase logic without a GPU; parity of
the real kernels is tests/test_gpu_parity.py (`-m gpu`).  Nothing here is a product path."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_close, load_model, make_extra_models, random_inputs

EMU_SRC = os.path.join(ROOT, "tests", "cpp", "coop_emu.cu")
EMU_LIB = os.path.join(ROOT, "tests", "cpp", "libcoop_emu.so")
CSRC = os.path.join(ROOT, "pinocchio_b200", "csrc")


def _emu():
    deps = [EMU_SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(EMU_LIB) or os.path.getmtime(EMU_LIB) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-Xcompiler", "-fPIC,-pthread", "-ccbin", cxx, "-shared", "-o", EMU_LIB, EMU_SRC, "-lpthread"])
    L = ctypes.CDLL(EMU_LIB)
    L.emu_last_error.restype = ctypes.c_char_p
    return L


@pytest.fixture(scope="module")
def emu():
    return _emu()


def _models():
    extra = make_extra_models()
    names = ["manipulator", "humanoid", "simple_humanoid_ff", "talos_reduced_ff", "mixed", "double_ff", "unaligned", "humanoid_hands", "wheeled"]
    return names, extra


def _get(name, extra):
    return extra[name] if name in extra else load_model(name)


def _run(emu, algo, model, q, v, x):
    from pinocchio_b200 import _capi
    fm, keep = _capi.make_flat(model.flat())
    B, nv = q.shape[1], model.nv
    outs = [np.full((nv * nv, B), np.nan, order="F") for _ in range(3)] + [np.full((nv, B), np.nan, order="F")]
    ins = [np.asfortranarray(z, dtype=np.float64) for z in (q, v, x)]
    p = lambda z: z.ctypes.data_as(ctypes.c_void_p)
    st = emu.emu_derivatives(ctypes.c_int(algo), ctypes.byref(fm), p(ins[0]), p(ins[1]), p(ins[2]), p(outs[0]), p(outs[1]),
                             p(outs[2]), p(outs[3]), ctypes.c_int64(B), ctypes.c_int(0))
    assert st == 0, emu.emu_last_error().decode()
    return outs


@pytest.mark.parametrize("name", _models()[0])
def test_emulated_rnea_derivatives(emu, oracle_cls, name):
    model = _get(name, _models()[1])
    q, v, a = random_inputs(model, 3, 11)
    dq, dv, da, tau = _run(emu, 0, model, q, v, a)
    rdq, rdv, rda, rtau = oracle_cls(model).rnea_derivatives(q, v, a)
    s = max(1.0, np.abs(rdq).max())
    assert_close(tau, rtau, atol=1e-12 * s, what="tau")
    assert_close(dq, rdq, atol=1e-12 * s, what="dtau_dq")
    assert_close(dv, rdv, atol=1e-12 * s, what="dtau_dv")
    assert_close(da, rda, atol=1e-12 * s, what="dtau_da")


@pytest.mark.parametrize("name", _models()[0])
def test_emulated_aba_derivatives(emu, oracle_cls, name):
    model = _get(name, _models()[1])
    q, v, tau = random_inputs(model, 3, 13)
    dq, dv, dt, ddq = _run(emu, 1, model, q, v, tau)
    rdq, rdv, rdt, rddq = oracle_cls(model).aba_derivatives(q, v, tau)
    for got, ref, what in ((ddq, rddq, "ddq"), (dt, rdt, "Minv"), (dq, rdq, "ddq_dq"), (dv, rdv, "ddq_dv")):
        assert np.isfinite(got).all(), what
        assert_close(got, ref, atol=1e-12 + 1e-10 * np.abs(ref).max(), what=what)


@pytest.mark.parametrize("name", _models()[0])
def test_emulated_minverse(emu, oracle_cls, name):
    model = _get(name, _models()[1])
    q, v, tau = random_inputs(model, 2, 17)
    _, _, Minv, _ = _run(emu, 2, model, q, v, tau)
    ref = oracle_cls(model).minverse(q)
    assert_close(Minv, ref, atol=1e-12 + 1e-10 * np.abs(ref).max(), what="Minv (upper)")
    nv = model.nv
    for b in range(q.shape[1]):
        assert not np.tril(Minv[:, b].reshape(nv, nv, order="F"), -1).any()


@pytest.mark.parametrize("blocked", [False, True])
@pytest.mark.parametrize("name", _models()[0] + ["humanoid_random"])
def test_emulated_minverse_cholesky(emu, oracle_cls, name, blocked):
    """computeMinverse as the engine runs it by default (minv_chol.cuh): Cholesky of crba's matrix, two triangular substitutions,
    G lanes per configuration — against the oracle's articulated-body Minv (aba.hxx:613-902); the lower triangle stays zero and
    the strictly-lower part of the input (garbage here) is never read."""
    model = _get(name, _models()[1])
    orc = oracle_cls(model)
    q, _, _ = random_inputs(model, 2, 23)
    nv = model.nv
    M = np.asfortranarray(orc.crba(q, world=True))
    for b in range(q.shape[1]):
        A = M[:, b].reshape(nv, nv, order="F")
        A[np.tril_indices(nv, -1)] = np.nan  # only the upper triangle may be read
        M[:, b] = A.reshape(-1, order="F")
    out = np.full((nv * nv, q.shape[1]), np.nan, order="F")
    p = lambda z: z.ctypes.data_as(ctypes.c_void_p)
    run = emu.emu_minv_chol_blocked_run if blocked else emu.emu_minv_chol_run  # 4 x 4 blocks, 32 lanes / the simple loops
    assert run(ctypes.c_int(nv), p(M), p(out), ctypes.c_int64(q.shape[1])) == 0
    ref = orc.minverse(q)
    assert_close(out, ref, atol=1e-12 + 1e-11 * np.abs(ref).max(), what="Minv (Cholesky)")
    for b in range(q.shape[1]):
        assert not np.tril(out[:, b].reshape(nv, nv, order="F"), -1).any()


@pytest.mark.parametrize("name", _models()[0])
def test_emulated_small_batch_rnea_and_aba(emu, oracle_cls, name):
    """The small-batch (cooperative) kernels of rneaInParallel / abaInParallel; aba(rnea(a)) == a closes the loop
    (unittest/aba.cpp:143-154)."""
    model = _get(name, _models()[1])
    orc = oracle_cls(model)
    q, v, a = random_inputs(model, 3, 19)
    tau = _run(emu, 3, model, q, v, a)[3]
    assert_close(tau, orc.rnea(q, v, a), atol=1e-12 * max(1.0, np.abs(tau).max()), what="rnea (coop)")
    ddq = _run(emu, 4, model, q, v, tau)[3]
    ref = orc.aba(q, v, tau)
    assert_close(ddq, ref, atol=1e-12 + 1e-10 * np.abs(ref).max(), what="aba (coop)")
    assert_close(ddq, a, atol=1e-12 + 1e-9 * np.abs(a).max(), what="aba(rnea(a)) == a")
