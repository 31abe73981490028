"""CPU tests of the per-model code generator (pinocchio_b200/csrc/codegen*): the generated straight-line program, emitted in
its host-callable variant and compiled with g++, against the oracle.  The device variant of the same text is what
brbd_pool_specialize compiles with NVRTC (GPU parity: tests/test_gpu_large.py::test_specialized_kernels).  The host variant
exists for these tests only; nothing in the product executes it.
Reference being mirrored: the code-generation unit tests, unittest/cppadcg-algo.cpp (generated code == algorithm)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import MODEL_NAMES, load_model, make_extra_models, random_inputs

ALL = MODEL_NAMES + ["mixed", "double_ff", "unaligned", "humanoid_hands", "wheeled"]
EXTRA = make_extra_models()


def build_host(model, algo, explicit_slots, tmp, **kw):
    from pinocchio_b200.codegen import codegen_source
    src, info = codegen_source(model, algo, explicit_slots=explicit_slots, host=True, **kw)
    cpp = os.path.join(tmp, f"gen_{algo}_{int(explicit_slots)}_{'_'.join(str(v) for v in kw.values())}.cpp")
    with open(cpp, "w") as fh:
        fh.write(src)
    so = cpp[:-4] + ".so"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O1", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, cpp, "-lm"])
    return getattr(ctypes.CDLL(so), f"brbd_gen_{algo}_host"), info


def run_host(fn, info, model, q, v, x, nout=None):
    B = q.shape[1]
    nout = model.nv if nout is None else nout
    out = np.zeros((nout, B), order="F")
    rec, park = np.full(max(1, info["record_slots"]), np.nan), np.full(max(1, info["park_slots"]), np.nan)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for i in range(B):
        qi, vi, xi = (np.ascontiguousarray(a[:, i]) for a in (q, v, x))
        oi = np.full(nout, np.nan)
        rec[:] = np.nan  # a read of a slot nobody wrote would poison the result
        park[:] = np.nan
        fn(P(qi), P(vi), P(xi), P(oi), P(rec), P(park))
        out[:, i] = oi
    return out


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("explicit_slots", [False, True])
def test_generated_program_matches_the_oracle(oracle_cls, name, explicit_slots):
    model = EXTRA[name] if name in EXTRA else load_model(name)
    orc = oracle_cls(model)
    q, v, x = random_inputs(model, 6, 11)
    with tempfile.TemporaryDirectory() as tmp:
        for algo, ref in (("rnea", orc.rnea(q, v, x)), ("aba", orc.aba(q, v, x))):
            fn, info = build_host(model, algo, explicit_slots, tmp)
            got = run_host(fn, info, model, q, v, x)
            assert np.isfinite(got).all(), (name, algo)
            assert np.abs(got - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), (name, algo, np.abs(got - ref).max())
            if explicit_slots and algo == "aba" and model.njoints > 3:
                assert info["park_slots"] > 0
        if not explicit_slots:
            fn, info = build_host(model, "crba", False, tmp)
            got = run_host(fn, info, model, q, v, x, nout=model.nv * model.nv)
            ref = orc.crba(q, world=True)
            assert np.isfinite(got).all(), (name, "crba")  # every entry written, zeros outside the tree sparsity
            assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (name, "crba")
            assert not got[ref == 0].any() or np.abs(got[ref == 0]).max() < 1e-13


def crba_pattern(model):
    """The structural pattern through the C ABI (brbd_model_crba_pattern)."""
    from pinocchio_b200 import _capi
    L = _capi.lib()
    fm, keep = _capi.make_flat(model.flat())
    h = ctypes.c_void_p()
    _capi.check(L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)))
    try:
        n = ctypes.c_int64()
        _capi.check(L.brbd_model_crba_pattern(h, None, None, 0, ctypes.byref(n)))
        rows, cols = np.zeros(n.value, dtype=np.int32), np.zeros(n.value, dtype=np.int32)
        _capi.check(L.brbd_model_crba_pattern(h, rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)))
        with pytest.raises(_capi.EngineError):
            _capi.check(L.brbd_model_crba_pattern(h, rows.ctypes.data_as(ctypes.c_void_p), None, n.value - 1, ctypes.byref(n)))
        return rows, cols
    finally:
        L.brbd_model_destroy(h)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("group", [1, 3, 31])
def test_generated_crba_compact_staging(oracle_cls, name, group):
    """CRBA with compact staging: the staging row holds the entries of the structural pattern only; the flush expands it through
    the position table (dense result) or writes it as it is (packed result, brbd_crba_packed_batch).  Both against the oracle,
    and the pattern covers every non-zero of the oracle's matrix whatever the column grouping."""
    model = EXTRA[name] if name in EXTRA else load_model(name)
    orc = oracle_cls(model)
    q, v, x = random_inputs(model, 4, 5)
    rows, cols = crba_pattern(model)
    nv, nnz = model.nv, len(rows)
    key = cols.astype(np.int64) * nv + rows
    assert (np.diff(key) > 0).all()  # column-major, no duplicates
    from conftest import structural_mask
    assert np.array_equal(np.flatnonzero(structural_mask(model)), key)  # = the tree sparsity of crba.hxx:94-95
    ref = orc.crba(q, world=True)
    outside = np.ones(nv * nv, dtype=bool)
    outside[key] = False
    assert not ref[outside].any()  # the pattern covers the oracle's non-zeros
    with tempfile.TemporaryDirectory() as tmp:
        fn, info = build_host(model, "crba", False, tmp, crba_compact=True, crba_group=group)
        got = run_host(fn, info, model, q, v, x, nout=nv * nv + nnz)
    assert np.isfinite(got).all(), name  # every dense and packed entry written
    dense, packed = got[: nv * nv], got[nv * nv:]
    tol = 1e-12 * max(1.0, np.abs(ref).max())
    assert np.abs(dense - ref).max() <= tol, (name, np.abs(dense - ref).max())
    assert not dense[outside].any()  # exact zeros outside the pattern
    assert np.abs(packed - ref[key]).max() <= tol, name
    assert np.array_equal(packed, dense[key])  # the two output modes carry the same values


@pytest.mark.parametrize("name", ["simple_humanoid_ff", "talos_reduced_ff", "mixed", "double_ff", "humanoid_hands"])
@pytest.mark.parametrize("group,nbuf", [(1, 1), (3, 1), (8, 1), (3, 2), (2, 3)])
def test_generated_crba_bulk_staging(oracle_cls, name, group, nbuf):
    """The staging logic of the bulk-copy CRBA wrapper, emulated for one lane on the host: `nbuf` staging rows that persist from
    configuration to configuration (as a lane's rows do from round to round), groups of adjacent columns staged `sh` elements in
    with `sh` changing from call to call (the destination's misalignment), stale entries cleared at the shift they were written
    with.  Every configuration must come out as the oracle's matrix with exact zeros outside the sparsity — a stale entry of an
    earlier group or configuration shows up as a non-zero there."""
    from conftest import structural_mask
    model = EXTRA[name] if name in EXTRA else load_model(name)
    orc = oracle_cls(model)
    nv = model.nv
    q, v, x = random_inputs(model, 7, 13)
    ref = orc.crba(q, world=True)
    mask = structural_mask(model)
    with tempfile.TemporaryDirectory() as tmp:
        fn, info = build_host(model, "crba", False, tmp, crba_bulk=True, crba_group=group, crba_nbuf=nbuf)
        lib = fn._objects if hasattr(fn, "_objects") else None
        so = [f for f in os.listdir(tmp) if f.endswith(".so")][0]
        par = ctypes.c_int.in_dll(ctypes.CDLL(os.path.join(tmp, so)), "brbd_gen_crba_host_par")
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rec, park = np.zeros(1), np.zeros(1)
        for i in range(q.shape[1]):
            par.value = (3 * i + 1) % 2  # another misalignment every call
            qi = np.ascontiguousarray(q[:, i])
            oi = np.full(nv * nv, np.nan)
            fn(P(qi), P(qi), P(qi), P(oi), P(rec), P(park))
            assert np.isfinite(oi).all(), (name, i)
            assert np.abs(oi - ref[:, i]).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (name, group, nbuf, i)
            assert not oi[~mask].any(), (name, group, nbuf, i, "stale staging entries")


def test_constant_folding_shrinks_the_program(oracle_cls):
    """simple_humanoid's placements are identity rotations: the generated ABA executes far fewer operations than the
    algorithm's count on a general model (26 103, the oracle's counting scalar), and the model's constants never appear as loads."""
    from pinocchio_b200.codegen import codegen_source
    model = load_model("simple_humanoid_ff")
    _, info = codegen_source(model, "aba")
    orc = oracle_cls(model)
    q, v, x = random_inputs(model, 1, 3)
    counted = orc.count_flops("aba", q[:, 0], v[:, 0], x[:, 0])["flops"]
    assert info["adds"] + info["muls"] < 0.75 * counted
    assert info["sincos"] == 29 and info["record_slots"] > 0
    src, _ = codegen_source(model, "aba", explicit_slots=True, nt=448)
    assert "tcgen05.st" in src and "brbd_gen_aba_0" in src


def test_codegen_rejects_unknown_algorithm():
    from pinocchio_b200 import _capi
    from pinocchio_b200.codegen import codegen_source, ALGOS
    ALGOS["bogus"] = 7
    try:
        with pytest.raises(_capi.EngineError):
            codegen_source(load_model("manipulator"), "bogus")
    finally:
        del ALGOS["bogus"]
