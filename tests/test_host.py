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
