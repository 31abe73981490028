import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_DIR = os.path.join(ROOT, "tests", "golden", "models")
MODEL_NAMES = ["manipulator", "humanoid", "humanoid_random", "simple_humanoid_ff", "talos_reduced_ff"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_model(name):
    from pinocchio_b200.model import Model
    with open(os.path.join(MODEL_DIR, name + ".json")) as fh:
        return Model.from_json(fh.read())


def make_extra_models():
    """Small models covering the joint families the config models do not (prismatic, spherical, planar)."""
    from pinocchio_b200 import model as M
    rng = M._Rng(7)
    out = {}
    # planar base + prismatic + spherical + revolute branches
    m = M.Model()
    m.name = "mixed"

    def add(jt, parent, name):
        nqj = M.joint_nq(jt)
        idx = m.addJoint(parent, jt, rng.se3(), name, np.full(nqj, -1.0), np.full(nqj, 1.0))
        m.appendBodyToJoint(idx, rng.inertia(), M.SE3.Identity())
        return idx
    b = add(M.JOINT_PLANAR, 0, "planar_base")
    p1 = add(M.JOINT_PX, b, "px")
    s1 = add(M.JOINT_SPHERICAL, p1, "sph1")
    add(M.JOINT_RZ, s1, "rz_tip")
    p2 = add(M.JOINT_PY, b, "py")
    r2 = add(M.JOINT_RX, p2, "rx")
    add(M.JOINT_PZ, r2, "pz")
    add(M.JOINT_SPHERICAL, r2, "sph2")
    m.armature = np.abs(rng.sym(m.nv)) * 0.1
    out["mixed"] = m
    # two free-flyers in a chain (multi-dof joint that is NOT the root) + armature
    m2 = M.Model()
    m2.name = "double_ff"
    lo7, hi7 = np.array([-1.0] * 7), np.array([1.0] * 7)
    f1 = m2.addJoint(0, M.JOINT_FREEFLYER, rng.se3(), "ff1", lo7, hi7)
    m2.appendBodyToJoint(f1, rng.inertia())
    r = m2.addJoint(f1, M.JOINT_RY, rng.se3(), "ry", [-1.0], [1.0])
    m2.appendBodyToJoint(r, rng.inertia())
    f2 = m2.addJoint(r, M.JOINT_FREEFLYER, rng.se3(), "ff2", lo7, hi7)
    m2.appendBodyToJoint(f2, rng.inertia())
    t = m2.addJoint(f2, M.JOINT_RX, rng.se3(), "rx", [-1.0], [1.0])
    m2.appendBodyToJoint(t, rng.inertia())
    m2.armature = np.abs(rng.sym(m2.nv)) * 0.05
    out["double_ff"] = m2
    # joints about arbitrary axes (JointModelRevoluteUnaligned / PrismaticUnaligned): a chain with a branch, children
    # of every kind below an unaligned joint, axes including -z (the half-turn case of the re-framing) and +z
    m3 = M.Model()
    m3.name = "unaligned"

    def addu(jt, parent, name, axis=None):
        nqj = M.joint_nq(jt)
        idx = m3.addJoint(parent, jt, rng.se3(), name, np.full(nqj, -1.0), np.full(nqj, 1.0), axis=axis)
        m3.appendBodyToJoint(idx, rng.inertia(), M.SE3.Identity())
        return idx
    u1 = addu(M.JOINT_REVOLUTE_UNALIGNED, 0, "ru1", [0.3, -0.5, 0.8])
    u2 = addu(M.JOINT_PRISMATIC_UNALIGNED, u1, "pu1", [-0.7, 0.2, 0.1])
    u3 = addu(M.JOINT_REVOLUTE_UNALIGNED, u2, "ru2", [0.0, 0.0, -2.0])
    addu(M.JOINT_RY, u3, "ry_tip")
    s3 = addu(M.JOINT_SPHERICAL, u1, "sph")
    u4 = addu(M.JOINT_REVOLUTE_UNALIGNED, s3, "ru3", [1.0, 1.0, 1.0])
    addu(M.JOINT_PRISMATIC_UNALIGNED, u4, "pu2", [0.0, 0.0, 1.0])
    addu(M.JOINT_REVOLUTE_UNALIGNED, u4, "ru4", [0.0, 1e-9, 1.0])
    m3.armature = np.abs(rng.sym(m3.nv)) * 0.05
    out["unaligned"] = m3
    # a 58-dof humanoid with three-finger hands (54 joints, depth 14, three branching joints on one root path): beyond the
    # 48 joints / dofs of round 1, and beyond what some on-chip layouts hold, so the launch code must degrade to the generic
    # kernels instead of refusing the model
    m4 = M.Model()
    m4.name = "humanoid_hands"

    def addh(jt, parent, name):
        nqj = M.joint_nq(jt)
        idx = m4.addJoint(parent, jt, rng.se3(), name, np.full(nqj, -1.0), np.full(nqj, 1.0))
        m4.appendBodyToJoint(idx, rng.inertia(), M.SE3.Identity())
        return idx
    rev = [M.JOINT_RX, M.JOINT_RY, M.JOINT_RZ]
    root = addh(M.JOINT_FREEFLYER, 0, "root")
    for side in ("l", "r"):
        p = root
        for k in range(6):
            p = addh(rev[(k + 2) % 3], p, f"leg_{side}{k}")
    t = addh(M.JOINT_RZ, root, "torso0")
    t = addh(M.JOINT_RY, t, "torso1")
    for side in ("l", "r"):
        p = t
        for k in range(7):
            p = addh(rev[k % 3], p, f"arm_{side}{k}")
        for f in range(3):
            pf = p
            for k in range(4):
                pf = addh(rev[(k + f) % 3] if k else M.JOINT_RZ, pf, f"finger_{side}{f}{k}")
    out["humanoid_hands"] = m4
    # unbounded revolute joints (URDF "continuous": RUBX / RUBY / RUBZ / RevoluteUnboundedUnaligned, q = (cos, sin)): a wheeled
    # base — planar joint, two wheels, a turret on a continuous joint about an oblique axis carrying a 2-dof arm
    m5 = M.Model()
    m5.name = "wheeled"

    def addw(jt, parent, name, axis=None):
        nqj = M.joint_nq(jt)
        idx = m5.addJoint(parent, jt, rng.se3(), name, np.full(nqj, -1.0), np.full(nqj, 1.0), axis=axis)
        m5.appendBodyToJoint(idx, rng.inertia(), M.SE3.Identity())
        return idx
    base = addw(M.JOINT_PLANAR, 0, "base")
    addw(M.JOINT_RUBY, base, "wheel_l")
    addw(M.JOINT_RUBX, base, "wheel_r")
    tur = addw(M.JOINT_REVOLUTE_UNBOUNDED_UNALIGNED, base, "turret", [0.2, -0.3, 0.9])
    sh = addw(M.JOINT_RUBZ, tur, "shoulder")
    addw(M.JOINT_RY, sh, "elbow")
    m5.armature = np.abs(rng.sym(m5.nv)) * 0.05
    out["wheeled"] = m5
    return out


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


def random_inputs(model, B, seed):
    """(q, v, a) with the reference benchmark's distributions (benchmark/timings-parallel.cpp:48-60)."""
    from pinocchio_b200.joint_configuration import batched_random_configuration, batched_random_tangent
    q = batched_random_configuration(model, B, seed)
    v = batched_random_tangent(model, B, seed + 1)
    a = batched_random_tangent(model, B, seed + 2)
    return q, v, a


# worst element-wise errors seen by assert_close, keyed by the first word of `what` (the algorithm): printed at the end of
# the session so that the gap between "1e-10 relative / 1e-12 absolute" and what the kernels deliver is a number
ERROR_LOG = {}


def _record(what, err, expected):
    if not what or err.size == 0:
        return
    key = what.split()[0]
    k = np.unravel_index(np.argmax(err), err.shape)
    scale = np.abs(expected).max()
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(np.abs(expected) > 0, err / np.abs(expected), 0.0)
    rec = ERROR_LOG.setdefault(key, {"abs": 0.0, "abs_over_max": 0.0, "rel_elem": 0.0, "n": 0, "where": ""})
    rec["n"] += int(err.size)
    if err[k] > rec["abs"]:
        rec["abs"], rec["where"] = float(err[k]), what
    rec["abs_over_max"] = max(rec["abs_over_max"], float(err[k] / scale) if scale > 0 else 0.0)
    # element-wise relative error only where the reference entry is not itself a cancellation residue
    big = np.abs(expected) > 1e-6 * scale
    if big.any():
        rec["rel_elem"] = max(rec["rel_elem"], float(rel[big].max()))


def pytest_terminal_summary(terminalreporter):
    if not ERROR_LOG:
        return
    tr = terminalreporter
    tr.write_line("worst element-wise error per algorithm over every assert_close of this session "
                  "(abs | abs / max|reference| | relative on entries > 1e-6 max|reference|):")
    for key in sorted(ERROR_LOG):
        r = ERROR_LOG[key]
        tr.write_line(f"  {key:28s} abs {r['abs']:.3e}  abs/max {r['abs_over_max']:.3e}  rel {r['rel_elem']:.3e}  "
                      f"({r['n']} entries; worst in: {r['where']})")
    try:
        import json
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_errors.json"), "w") as fh:
            json.dump(ERROR_LOG, fh, indent=1)
    except OSError:
        pass


def assert_close(actual, expected, rtol=1e-10, atol=1e-12, what=""):
    """north_star tolerance: 1e-10 relative / 1e-12 absolute (element-wise, numpy allclose semantics)."""
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    assert np.isfinite(actual).all(), f"{what}: non-finite entries in the result"
    err = np.abs(actual - expected)
    _record(what, err, expected)
    tol = atol + rtol * np.abs(expected)
    bad = err > tol
    if bad.any():
        k = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: {int(bad.sum())} entries out of tolerance; worst at {k}: "
                             f"got {actual[k]!r} expected {expected[k]!r} (err {err[k]:.3e}, tol {tol[k]:.3e})")


@pytest.fixture(scope="session")
def oracle_cls():
    from oracle import Oracle, build_oracle
    build_oracle()
    return Oracle
