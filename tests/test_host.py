"""CPU-side tests: host model logic, fixtures, and that the C-ABI library loads and exports every
symbol include/pinocchio_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import MODEL_NAMES, ROOT, load_model, make_extra_models


def test_header_symbols_exported():
    from pinocchio_b200 import _capi
    header = open(os.path.join(ROOT, "include", "pinocchio_b200.h")).read()
    declared = set(re.findall(r"\b(brbd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = _capi.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert b"sm_100a" in L.brbd_version()


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when no CUDA device is usable."""
    import pinocchio_b200 as pb
    from pinocchio_b200 import _capi
    if _capi.lib().brbd_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(Exception) as ei:
        pb.ModelPool(load_model("manipulator"))
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's cpu legs may touch oracle/."""
    pkg = os.path.join(ROOT, "pinocchio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "oracle" not in text, (dirpath, f)


def test_model_validation_errors():
    from pinocchio_b200 import _capi
    L = _capi.lib()
    m = load_model("manipulator")
    flat = m.flat()
    bad = dict(flat)
    bad["joint_type"] = flat["joint_type"].copy()
    bad["joint_type"][3] = 17
    fm, keep = _capi.make_flat(bad)
    h = ctypes.c_void_p()
    assert L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)) == _capi.BRBD_EUNSUPPORTED_JOINT
    assert b"unsupported" in L.brbd_last_error_string()
    bad = dict(flat)
    bad["parents"] = flat["parents"].copy()
    bad["parents"][2] = 5
    fm, keep = _capi.make_flat(bad)
    assert L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)) == _capi.BRBD_ETOPOLOGY
    fm, keep = _capi.make_flat(flat)
    assert L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)) == _capi.BRBD_OK
    assert (L.brbd_model_nq(h), L.brbd_model_nv(h), L.brbd_model_njoints(h)) == (6, 6, 7)
    L.brbd_model_destroy(h)


def test_non_compact_tree_rejected():
    """CRBAChecker (crba.hxx:573-595): joints must be numbered depth-first."""
    from pinocchio_b200 import _capi, model as M
    m = M.Model()
    a = m.addJoint(0, M.JOINT_RX, M.SE3.Identity(), "a", [-1], [1])
    b = m.addJoint(0, M.JOINT_RY, M.SE3.Identity(), "b", [-1], [1])
    m.addJoint(a, M.JOINT_RZ, M.SE3.Identity(), "c", [-1], [1])  # child of a after b: not compact
    for j in range(1, 4):
        m.appendBodyToJoint(j, M.Inertia(1.0, [0.1, 0, 0], np.eye(3)))
    assert not m.is_compact()
    fm, keep = _capi.make_flat(m.flat())
    h = ctypes.c_void_p()
    assert _capi.lib().brbd_model_create(ctypes.byref(fm), ctypes.byref(h)) == _capi.BRBD_ETOPOLOGY


def test_fixtures_match_builders():
    """The committed JSON fixtures are what the builders produce (and survive a round trip)."""
    from pinocchio_b200 import model as M
    for name, build in (("manipulator", M.buildSampleModelManipulator), ("humanoid", M.buildSampleModelHumanoid),
                        ("humanoid_random", M.buildSampleModelHumanoidRandom)):
        a, b = load_model(name).flat(), build().flat()
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (name, k)


def test_urdf_rules_on_inline_model():
    """Fixed joints merge into the parent; children are visited in joint-name order; axis -> RX/RY/RZ/P*."""
    from pinocchio_b200 import model as M
    urdf = """<robot name="t">
      <link name="base"><inertial><mass value="2"/><origin xyz="0 0 0.1"/><inertia ixx="1" iyy="1" izz="1" ixy="0" ixz="0" iyz="0"/></inertial></link>
      <link name="l1"><inertial><mass value="1"/><origin xyz="0.1 0 0"/><inertia ixx="0.1" iyy="0.2" izz="0.3" ixy="0" ixz="0" iyz="0"/></inertial></link>
      <link name="tool"><inertial><mass value="0.5"/><origin xyz="0 0 0.2"/><inertia ixx="0.01" iyy="0.01" izz="0.01" ixy="0" ixz="0" iyz="0"/></inertial></link>
      <link name="l2"><inertial><mass value="1.5"/><origin xyz="0 0.1 0"/><inertia ixx="0.1" iyy="0.1" izz="0.1" ixy="0" ixz="0" iyz="0"/></inertial></link>
      <joint name="zz_first_in_file" type="revolute"><parent link="base"/><child link="l2"/><origin xyz="0 1 0"/><axis xyz="0 0 1"/><limit lower="-1" upper="1" effort="1" velocity="1"/></joint>
      <joint name="a_joint" type="prismatic"><parent link="base"/><child link="l1"/><origin xyz="1 0 0" rpy="0 0 1.5707963267948966"/><axis xyz="0 1 0"/><limit lower="-1" upper="1" effort="1" velocity="1"/></joint>
      <joint name="fix" type="fixed"><parent link="l1"/><child link="tool"/><origin xyz="0 0 0.5"/></joint>
    </robot>"""
    m = M.buildModelFromUrdf(urdf, M.JOINT_FREEFLYER)
    assert m.names == ["universe", "root_joint", "a_joint", "zz_first_in_file"]
    assert m.joint_types[1:] == [M.JOINT_FREEFLYER, M.JOINT_PY, M.JOINT_RZ]
    assert (m.nq, m.nv) == (9, 8)
    assert abs(m.inertias[2].mass - 1.5) < 1e-15  # l1 + tool merged
    assert np.allclose(m.inertias[2].lever, (1.0 * np.array([0.1, 0, 0]) + 0.5 * np.array([0, 0, 0.7])) / 1.5)
    assert np.allclose(m.jointPlacements[2].R, [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15)
    # any other axis -> JointModelRevoluteUnaligned(axis.normalized()) (parsers/urdf/model.hxx:437-480)
    mu = M.buildModelFromUrdf(urdf.replace('axis xyz="0 0 1"', 'axis xyz="0 0.3 0.4"'), M.JOINT_FREEFLYER)
    assert mu.joint_types[1:] == [M.JOINT_FREEFLYER, M.JOINT_PY, M.JOINT_REVOLUTE_UNALIGNED]
    assert np.allclose(mu.axes[3], [0, 0.6, 0.8], atol=1e-15) and (mu.nq, mu.nv) == (9, 8)


def test_inertia_algebra_against_dense_matrices():
    """unittest/spatial.cpp:515-631: X.I as a 6x6 equals X^-T I X^-1; I+I as matrices."""
    from pinocchio_b200 import model as M
    rng = M._Rng(3)
    for _ in range(5):
        X, Y1, Y2 = rng.se3(), rng.inertia(), rng.inertia()
        px = np.array([[0, -X.p[2], X.p[1]], [X.p[2], 0, -X.p[0]], [-X.p[1], X.p[0], 0]])
        A = np.zeros((6, 6))  # action matrix on motions (linear first)
        A[:3, :3], A[3:, 3:], A[:3, 3:] = X.R, X.R, px @ X.R
        Ainv = np.linalg.inv(A)
        assert np.allclose(Y1.se3Action(X).matrix(), Ainv.T @ Y1.matrix() @ Ainv, atol=1e-12)
        S = Y1.copy()
        S += Y2
        assert np.allclose(S.matrix(), Y1.matrix() + Y2.matrix(), atol=1e-12)


def test_random_configuration_stream():
    """randomConfiguration consumes libc rand() joint by joint (joint-configuration.hxx:177-185)."""
    from pinocchio_b200.joint_configuration import LibcRand, RAND_MAX, neutral, randomConfiguration
    m = load_model("humanoid_random")
    q1 = randomConfiguration(m, -1, 1, LibcRand(5))
    q2 = randomConfiguration(m, -1, 1, LibcRand(5))
    assert np.array_equal(q1, q2)
    assert abs(np.linalg.norm(q1[3:7]) - 1.0) < 1e-14
    r = LibcRand(5)
    assert q1[0] == -1.0 + (2.0 * r.rand()) / RAND_MAX
    assert neutral(m)[6] == 1.0


def test_integrate_is_consistent_with_exp():
    from pinocchio_b200.joint_configuration import integrate, neutral
    for m in list(make_extra_models().values()) + [load_model("humanoid_random")]:
        q0 = neutral(m)
        v = np.linspace(-0.3, 0.4, m.nv)
        q1 = integrate(m, q0, v)
        q2 = integrate(m, integrate(m, q0, 0.5 * v), 0.5 * v)  # one-parameter subgroup from the neutral element
        assert np.allclose(q1, q2, atol=1e-12)


def test_pool_argument_marshalling_without_gpu():
    """_describe resolves pointer / leading dimension exactly like Eigen's data() / outerStride()."""
    from pinocchio_b200.pool import _describe
    a = np.asfortranarray(np.arange(12.0).reshape(3, 4))
    d = _describe(a, 3, "x")
    assert (d.ld, d.cols, d.device) == (3, 4, False)
    big = np.zeros((5, 4), order="F")
    d = _describe(big[:3], 3, "x")
    assert d.ld == 5
    with pytest.raises(ValueError):
        _describe(a, 4, "x")
    c = np.ascontiguousarray(a)  # row-major input is copied to column-major
    assert _describe(c, 3, "x").ld == 3
    with pytest.raises(ValueError):
        _describe(c, 3, "x", out=True)


WHEELED_URDF = """<?xml version="1.0"?>
<robot name="cart">
  <link name="base"><inertial><origin xyz="0 0 0.1"/><mass value="5"/><inertia ixx="0.1" ixy="0" ixz="0" iyy="0.2" iyz="0" izz="0.3"/></inertial></link>
  <link name="wheel_l"><inertial><mass value="1"/><inertia ixx="0.01" ixy="0" ixz="0" iyy="0.02" iyz="0" izz="0.01"/></inertial></link>
  <link name="wheel_r"><inertial><mass value="1"/><inertia ixx="0.01" ixy="0" ixz="0" iyy="0.02" iyz="0" izz="0.01"/></inertial></link>
  <link name="turret"><inertial><origin xyz="0.05 0 0.2"/><mass value="2"/><inertia ixx="0.05" ixy="0.001" ixz="0" iyy="0.05" iyz="0" izz="0.02"/></inertial></link>
  <joint name="jl" type="continuous"><parent link="base"/><child link="wheel_l"/><origin xyz="0 0.3 0" rpy="0 0 0"/><axis xyz="0 1 0"/></joint>
  <joint name="jr" type="continuous"><parent link="base"/><child link="wheel_r"/><origin xyz="0 -0.3 0"/><axis xyz="0 1 0"/></joint>
  <joint name="jt" type="continuous"><parent link="base"/><child link="turret"/><origin xyz="0.1 0 0.2"/><axis xyz="0.1 0.2 1"/></joint>
</robot>
"""


def test_urdf_continuous_joints_are_unbounded_revolute(oracle_cls):
    """parsers/urdf/model.hxx:269-273: CONTINUOUS -> JointModelRUBX / RUBY / RUBZ / RevoluteUnboundedUnaligned (nq = 2, q = (cos, sin),
    joint-revolute-unbounded.hpp:154-162).  Dynamics equal those of the bounded revolute joint at the same angle."""
    from pinocchio_b200 import model as M
    from pinocchio_b200.joint_configuration import neutral, randomConfiguration
    m = M.buildModelFromUrdf(WHEELED_URDF, root_joint=M.JOINT_FREEFLYER)
    assert [m.joint_types[m.getJointId(n)] for n in ("jl", "jr", "jt")] == [M.JOINT_RUBY, M.JOINT_RUBY, M.JOINT_REVOLUTE_UNBOUNDED_UNALIGNED]
    assert (m.nq, m.nv) == (7 + 2 * 3, 6 + 3)
    q0 = neutral(m)
    assert np.allclose(q0[7:], [1, 0, 1, 0, 1, 0])
    # the same robot with `revolute` joints: identical results at q_bounded = angle, q_unbounded = (cos, sin)(angle)
    mb = M.buildModelFromUrdf(WHEELED_URDF.replace('type="continuous"', 'type="revolute"'), root_joint=M.JOINT_FREEFLYER)
    rng = np.random.default_rng(3)
    qb = np.concatenate([rng.uniform(-1, 1, 3), [0.1, -0.2, 0.3, 0.0], rng.uniform(-3, 3, 3)])
    qb[3:7] /= np.linalg.norm(qb[3:7]) if np.linalg.norm(qb[3:7]) > 0 else 1.0
    qb[6] = np.sqrt(max(0.0, 1.0 - qb[3] ** 2 - qb[4] ** 2 - qb[5] ** 2))
    qu = np.concatenate([qb[:7], np.column_stack([np.cos(qb[7:]), np.sin(qb[7:])]).ravel()])
    v, a = rng.uniform(-1, 1, m.nv), rng.uniform(-1, 1, m.nv)
    ou, ob = oracle_cls(m), oracle_cls(mb)
    assert np.allclose(ou.rnea(qu, v, a), ob.rnea(qb, v, a), rtol=1e-12, atol=1e-12)
    assert np.allclose(ou.aba(qu, v, a), ob.aba(qb, v, a), rtol=1e-10, atol=1e-12)
    assert np.allclose(ou.crba(qu, world=True), ob.crba(qb, world=True), rtol=1e-12, atol=1e-12)
    q = randomConfiguration(m, np.full(m.nq, -1.0), np.full(m.nq, 1.0))
    assert np.allclose(q[7::2] ** 2 + q[8::2] ** 2, 1.0)
    with pytest.raises(ValueError, match="mimic"):
        M.buildModelFromUrdf(WHEELED_URDF.replace('<axis xyz="0 1 0"/></joint>', '<axis xyz="0 1 0"/><mimic joint="jt"/></joint>', 1))


def test_c_abi_accepts_unbounded_joint_tags():
    """brbd_model_create: tags 11..14, nq = 2 per joint; a wrong nq is refused."""
    from pinocchio_b200 import _capi
    import ctypes
    m = make_extra_models()["wheeled"]
    fm, keep = _capi.make_flat(m.flat())
    h = ctypes.c_void_p()
    L = _capi.lib()
    assert L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)) == _capi.BRBD_OK
    assert (L.brbd_model_nq(h), L.brbd_model_nv(h)) == (m.nq, m.nv) == (4 + 2 * 4 + 1, 3 + 5)
    L.brbd_model_destroy(h)
    fm.nq = m.nq - 1
    assert L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)) == _capi.BRBD_EINVAL


def _flat_from_c_urdf(text_or_path, root):
    """brbd_model_from_urdf + brbd_model_get_flat -> dict of numpy arrays."""
    import ctypes
    from pinocchio_b200 import _capi
    L = _capi.lib()
    h = ctypes.c_void_p()
    _capi.check(L.brbd_model_from_urdf(text_or_path.encode(), root, ctypes.byref(h)))
    fm = _capi.FlatModel()
    _capi.check(L.brbd_model_get_flat(h, ctypes.byref(fm)))
    n, nv = fm.njoints, fm.nv
    a = lambda ptr, k, dt: np.ctypeslib.as_array(ptr, shape=(k,)).astype(dt).copy()
    out = {"njoints": n, "nq": fm.nq, "nv": nv, "parents": a(fm.parents, n, np.int32), "joint_type": a(fm.joint_type, n, np.int32),
           "idx_q": a(fm.idx_q, n, np.int32), "idx_v": a(fm.idx_v, n, np.int32), "placement": a(fm.placement, 12 * n, np.float64).reshape(n, 12),
           "inertia": a(fm.inertia, 10 * n, np.float64).reshape(n, 10), "axis": a(fm.axis, 3 * n, np.float64).reshape(n, 3)}
    L.brbd_model_destroy(h)
    return out


def test_cpp_urdf_loader_matches_the_python_one():
    """brbd_model_from_urdf (urdf_loader.cpp) against pinocchio_b200.model.buildModelFromUrdf: same joints in the same order,
    same placements and merged inertias — on the wheeled cart (continuous + unaligned joints), on the reference's two robot
    files when they are present (models/simple_humanoid.urdf, talos_reduced.urdf: unittest/urdf.cpp:85,252), fixed base and
    free-flyer root."""
    from pinocchio_b200 import model as M
    from pinocchio_b200 import _capi
    cases = [(WHEELED_URDF, M.JOINT_FREEFLYER), (WHEELED_URDF, None)]
    for f in ("/root/reference/models/simple_humanoid.urdf",
              "/root/reference/models/example-robot-data/robots/talos_data/robots/talos_reduced.urdf"):
        if os.path.exists(f):
            cases.append((f, M.JOINT_FREEFLYER))
    for src, root in cases:
        py = M.buildModelFromUrdf(src, root_joint=root).flat()
        c = _flat_from_c_urdf(src, -1 if root is None else root)
        assert (c["njoints"], c["nq"], c["nv"]) == (py["njoints"], py["nq"], py["nv"]), src[:60]
        for k in ("parents", "joint_type", "idx_q", "idx_v"):
            assert np.array_equal(c[k], py[k]), (src[:60], k)
        assert np.allclose(c["placement"], py["placement"], rtol=0, atol=1e-15)
        # joint 0 (the universe) carries the fixed-base root link in the Python mirror only; it never enters an algorithm
        assert np.allclose(c["inertia"][1:], py["inertia"][1:], rtol=1e-14, atol=1e-15)
        assert np.allclose(c["axis"], py["axis"].reshape(-1, 3), atol=1e-15)
    if any(s.startswith("/root/reference/models/simple") for s, _ in cases):
        c = _flat_from_c_urdf("/root/reference/models/simple_humanoid.urdf", M.JOINT_FREEFLYER)
        assert (c["nq"], c["njoints"]) == (36, 31)  # unittest/urdf.cpp:85,252
    import ctypes
    h = ctypes.c_void_p()
    L = _capi.lib()
    assert L.brbd_model_from_urdf(b"<robot><link name='a'/><link name='b'/></robot>", -1, ctypes.byref(h)) == _capi.BRBD_EINVAL  # two roots
    bad = WHEELED_URDF.replace('<axis xyz="0 1 0"/></joint>', '<axis xyz="0 1 0"/><mimic joint="jt"/></joint>', 1)
    assert L.brbd_model_from_urdf(bad.encode(), -1, ctypes.byref(h)) == _capi.BRBD_EUNSUPPORTED_JOINT
    assert b"mimic" in L.brbd_last_error_string()
    assert L.brbd_model_from_urdf(b"/no/such/file.urdf", -1, ctypes.byref(h)) == _capi.BRBD_EINVAL


@pytest.mark.parametrize("name", ["simple_humanoid_ff", "talos_reduced_ff", "manipulator"])
def test_crba_expand_packed_on_the_host(name):
    """brbd_crba_expand_packed (csrc/host_expand.cpp): packed pattern entries -> dense column-major matrices, by host threads with
    non-temporal stores.  A copy, checked against a numpy scatter: dense and padded leading dimensions (an odd nv * nv puts every
    other matrix 8 bytes off a 16-byte boundary), a block that starts off a 16-byte boundary, one and several threads, FP32, a
    canary around the destination."""
    import ctypes
    from pinocchio_b200 import _capi
    from conftest import load_model
    model = load_model(name)
    L = _capi.lib()
    fm, keep = _capi.make_flat(model.flat())
    h = ctypes.c_void_p()
    _capi.check(L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)))
    try:
        n = ctypes.c_int64()
        _capi.check(L.brbd_model_crba_pattern(h, None, None, 0, ctypes.byref(n)))
        rows, cols = np.zeros(n.value, dtype=np.int32), np.zeros(n.value, dtype=np.int32)
        _capi.check(L.brbd_model_crba_pattern(h, rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)))
        nv, nnz = model.nv, n.value
        nn = nv * nv
        key = cols.astype(np.int64) * nv + rows
        rng = np.random.default_rng(3)
        B = 257
        for dt, flag in ((np.float64, _capi.BRBD_FP64), (np.float32, _capi.BRBD_FP32)):
            for padP, padM, shift, threads in ((0, 0, 0, 1), (3, 0, 1, 4), (0, 3, 0, 3), (1, 2, 1, 8)):
                Pbuf = rng.standard_normal((nnz + padP) * B).astype(dt)
                P = Pbuf.reshape(B, nnz + padP)[:, :nnz]  # configuration per row = column-major (nnz x B) with ld = nnz + padP
                flat = np.full((B + 1) * (nn + padM) + 4, -7.0, dtype=dt)
                ref = np.zeros((B, nn), dtype=dt)
                ref[:, key] = P
                dst = flat[shift:]
                _capi.check(L.brbd_crba_expand_packed(h, Pbuf.ctypes.data_as(ctypes.c_void_p), nnz + padP, dst.ctypes.data_as(ctypes.c_void_p),
                                                      nn + padM, B, threads, flag))
                got = flat[shift:shift + B * (nn + padM)].reshape(B, nn + padM)
                assert np.array_equal(got[:, :nn], ref), (name, dt, padP, padM, shift, threads)
                assert (got[:, nn:] == -7.0).all() and (flat[:shift] == -7.0).all() and (flat[shift + B * (nn + padM):] == -7.0).all()
        with pytest.raises(_capi.EngineError):
            _capi.check(L.brbd_crba_expand_packed(h, Pbuf.ctypes.data_as(ctypes.c_void_p), nnz - 1, flat.ctypes.data_as(ctypes.c_void_p), nn, B, 1, 0))
    finally:
        L.brbd_model_destroy(h)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU: one JSON line with the keys of the
    bench contract, the metric / unit / config of the GPU arm, `impl: reference`, an `e2e` that repeats the value with no copies."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["metric"] == "batched dynamics evals/sec (aba + crba)" and line["dtype"] == "f64" and line["n_gpus"] == 1
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert "65536" in line["config"]["workload"] and "simple_humanoid" in line["config"]["workload"]
