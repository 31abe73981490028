"""Pins the CPU oracle with the reference's own identity tests.

The reference ships no golden vectors for this path (unittest/rnea.cpp:5-9: "numerical values are
not cross validated in any way"); every test of RNEA / ABA / CRBA / derivatives there is an identity.
Each test below re-runs one of them against oracle/ and cites the reference test it restates.
"""
import numpy as np
import pytest

from conftest import MODEL_NAMES, load_model, make_extra_models, random_inputs

ALL = MODEL_NAMES + ["mixed", "double_ff", "unaligned", "humanoid_hands", "wheeled"]
EXTRA = make_extra_models()


def get_model(name):
    return EXTRA[name] if name in EXTRA else load_model(name)


def is_approx(a, b, prec=1e-12):
    """Eigen isApprox: ||a-b|| <= prec * min(||a||, ||b||)."""
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    return np.linalg.norm(a - b) <= prec * min(np.linalg.norm(a), np.linalg.norm(b))


def mat(col, nv):
    return np.asarray(col).reshape(nv, nv, order="F")


def sym_from_upper(M):
    return np.triu(M) + np.triu(M, 1).T


@pytest.fixture(scope="module")
def make(oracle_cls):
    cache = {}

    def get(name):
        if name not in cache:
            m = get_model(name)
            cache[name] = (m, oracle_cls(m))
        return cache[name]
    return get


def test_known_dimensions():
    """unittest/sample-models.cpp:43-44,70-71,92-93; unittest/urdf.cpp:85,252."""
    dims = {"manipulator": (6, 6, 7), "humanoid": (35, 34, 30), "humanoid_random": (33, 32, 28),
            "simple_humanoid_ff": (36, 35, 31), "talos_reduced_ff": (39, 38, 34)}
    for name, (nq, nv, nj) in dims.items():
        m = load_model(name)
        assert (m.nq, m.nv, m.njoints) == (nq, nv, nj), name


@pytest.mark.parametrize("name", ALL)
def test_aba_inverts_rnea(make, name):
    """unittest/aba.cpp:124-184: aba(q, v, rnea(q, v, a)) == a (WORLD convention), 1e-12."""
    m, o = make(name)
    q, v, a = random_inputs(m, 8, 1)
    tau = o.rnea(q, v, a)
    a2 = o.aba(q, v, tau)
    for i in range(8):
        assert is_approx(a2[:, i], a[:, i], 1e-12), (name, i)  # measured worst: talos 9.2e-13 (cond(M) = 8.7e4)


@pytest.mark.parametrize("name", ALL)
def test_crba_vs_rnea_columns(make, name):
    """unittest/crba.cpp:92-129: M[:, i] == rnea(q, 0, e_i) - rnea(q, 0, 0), 1e-12."""
    m, o = make(name)
    q, _, _ = random_inputs(m, 1, 2)
    nv = m.nv
    M = sym_from_upper(mat(o.crba(q)[:, 0], nv))
    z = np.zeros(nv)
    bias = o.rnea(q[:, 0], z, z)[:, 0]
    for i in range(nv):
        e = np.zeros(nv)
        e[i] = 1.0
        col = o.rnea(q[:, 0], z, e)[:, 0] - bias
        assert np.linalg.norm(col - M[:, i]) <= 1e-12 * max(1.0, np.linalg.norm(col)), (name, i)


@pytest.mark.parametrize("name", ALL)
def test_crba_local_equals_world(make, name):
    """unittest/crba.cpp:172-208."""
    m, o = make(name)
    q, _, _ = random_inputs(m, 4, 3)
    assert is_approx(o.crba(q, world=False), o.crba(q, world=True), 1e-12)


@pytest.mark.parametrize("name", ALL)
def test_equation_of_motion(make, name):
    """unittest/aba.cpp:229-263: M a + nle == rnea(q, v, a); aba(M a + nle) == a."""
    m, o = make(name)
    q, v, a = random_inputs(m, 2, 4)
    for i in range(2):
        M = sym_from_upper(mat(o.crba(q[:, i], world=True)[:, 0], m.nv))
        nle = o.rnea(q[:, i], v[:, i], np.zeros(m.nv))[:, 0]  # unittest/rnea.cpp:79-134
        tau = M @ a[:, i] + nle
        assert is_approx(tau, o.rnea(q[:, i], v[:, i], a[:, i])[:, 0], 1e-12)
        assert is_approx(o.aba(q[:, i], v[:, i], tau)[:, 0], a[:, i], 1e-12)


def test_armature(make, oracle_cls):
    """unittest/rnea.cpp:179-207 and unittest/crba.cpp:210-235: rotor inertia adds armature*a / diag(armature)."""
    m = load_model("humanoid_random")
    o0 = oracle_cls(m)
    m2 = load_model("humanoid_random")
    m2.armature = np.linspace(0.1, 1.0, m2.nv)
    o1 = oracle_cls(m2)
    q, v, a = random_inputs(m, 3, 5)
    assert is_approx(o1.rnea(q, v, a), o0.rnea(q, v, a) + m2.armature[:, None] * a, 1e-12)
    for i in range(3):
        M0, M1 = mat(o0.crba(q[:, i])[:, 0], m.nv), mat(o1.crba(q[:, i])[:, 0], m.nv)
        assert is_approx(M1, M0 + np.diag(m2.armature), 1e-12)
    tau = o1.rnea(q, v, a)
    assert is_approx(o1.aba(q, v, tau), a, 1e-12)  # unittest/aba.cpp:367-...


@pytest.mark.parametrize("name", ALL)
def test_rnea_derivatives_vs_finite_differences(make, name):
    """unittest/rnea-derivatives.cpp:113-313: forward differences, alpha = 1e-8, tolerance sqrt(alpha);
    dtau_da == crba (:294-299); data.tau == rnea."""
    from pinocchio_b200.joint_configuration import integrate
    m, o = make(name)
    q, v, a = random_inputs(m, 1, 6)
    q, v, a = q[:, 0], v[:, 0], a[:, 0]
    nv = m.nv
    dq, dv, da, tau = o.rnea_derivatives(q, v, a)
    tau0 = o.rnea(q, v, a)[:, 0]
    assert is_approx(tau[:, 0], tau0, 1e-12)
    assert is_approx(mat(da[:, 0], nv), mat(o.crba(q, world=True)[:, 0], nv), 1e-12)
    alpha = 1e-8
    fdq, fdv = np.zeros((nv, nv)), np.zeros((nv, nv))
    for k in range(nv):
        e = np.zeros(nv)
        e[k] = alpha
        fdq[:, k] = (o.rnea(integrate(m, q, e), v, a)[:, 0] - tau0) / alpha
        fdv[:, k] = (o.rnea(q, v + e, a)[:, 0] - tau0) / alpha
    assert is_approx(mat(dq[:, 0], nv), fdq, np.sqrt(alpha))
    assert is_approx(mat(dv[:, 0], nv), fdv, np.sqrt(alpha))


@pytest.mark.parametrize("name", ALL)
def test_aba_derivatives(make, name):
    """unittest/aba-derivatives.cpp:23-155: ddq_dtau == Minv == inv(crba); dq/dv == -Minv * dtau_d{q,v} of
    computeRNEADerivatives(q, v, aba(...)) (:107-108); partials vs finite differences at sqrt(alpha)."""
    from pinocchio_b200.joint_configuration import integrate
    m, o = make(name)
    q, v, tau = random_inputs(m, 1, 7)
    q, v, tau = q[:, 0], v[:, 0], tau[:, 0]
    nv = m.nv
    dq, dv, dtau, ddq = o.aba_derivatives(q, v, tau)
    a0 = o.aba(q, v, tau)[:, 0]
    assert is_approx(ddq[:, 0], a0, 1e-12)
    Minv = mat(dtau[:, 0], nv)
    M = sym_from_upper(mat(o.crba(q, world=True)[:, 0], nv))
    assert is_approx(Minv, np.linalg.inv(M), 1e-12)  # unittest/aba.cpp:265-340 (measured <= 1.3e-14)
    rdq, rdv, _, _ = o.rnea_derivatives(q, v, a0)
    assert is_approx(mat(dq[:, 0], nv), -Minv @ mat(rdq[:, 0], nv), 1e-12)  # unittest/aba-derivatives.cpp:107-108
    assert is_approx(mat(dv[:, 0], nv), -Minv @ mat(rdv[:, 0], nv), 1e-12)
    alpha = 1e-8
    fdq, fdv, fdt = np.zeros((nv, nv)), np.zeros((nv, nv)), np.zeros((nv, nv))
    for k in range(nv):
        e = np.zeros(nv)
        e[k] = alpha
        fdq[:, k] = (o.aba(integrate(m, q, e), v, tau)[:, 0] - a0) / alpha
        fdv[:, k] = (o.aba(q, v + e, tau)[:, 0] - a0) / alpha
        fdt[:, k] = (o.aba(q, v, tau + e)[:, 0] - a0) / alpha
    assert is_approx(mat(dq[:, 0], nv), fdq, np.sqrt(alpha))  # unittest/aba-derivatives.cpp:116-154 (measured <= 3.6e-6)
    assert is_approx(mat(dv[:, 0], nv), fdv, np.sqrt(alpha))
    assert is_approx(Minv, fdt, np.sqrt(alpha))


@pytest.mark.parametrize("name", ["manipulator", "humanoid_random", "talos_reduced_ff", "mixed"])
def test_double_vs_long_double(make, name):
    """Independent guard (SURVEY §7 hard part 3): the double oracle agrees with its 80-bit instantiation."""
    m, o = make(name)
    q, v, a = random_inputs(m, 4, 8)
    assert is_approx(o.rnea(q, v, a), o.rnea(q, v, a, long_double=True), 1e-12)
    tau = o.rnea(q, v, a)
    assert is_approx(o.aba(q, v, tau), o.aba(q, v, tau, long_double=True), 1e-10)
    assert is_approx(o.crba(q, world=True), o.crba(q, world=True, long_double=True), 1e-12)


@pytest.mark.parametrize("name", ALL)
def test_double_vs_float128(make, name):
    """Second independent guard: the oracle instantiated with IEEE binary128 (libquadmath, 113-bit mantissa).
    The double-precision products agree with it at 1e-13; the solves at 1e-12 (their error grows with cond(M): talos 8.7e4)."""
    m, o = make(name)
    q, v, a = random_inputs(m, 4, 8)
    assert is_approx(o.rnea(q, v, a), o.rnea(q, v, a, long_double=2), 1e-13)
    assert is_approx(o.crba(q, world=True), o.crba(q, world=True, long_double=2), 1e-13)
    tau = o.rnea(q, v, a)
    assert is_approx(o.aba(q, v, tau), o.aba(q, v, tau, long_double=2), 1e-12)
    for d, dq_ in zip(o.rnea_derivatives(q, v, a), o.rnea_derivatives(q, v, a, long_double=2)):
        assert is_approx(d, dq_, 1e-13)
    for d, dq_ in zip(o.aba_derivatives(q, v, tau), o.aba_derivatives(q, v, tau, long_double=2)):
        assert is_approx(d, dq_, 1e-11)


@pytest.mark.parametrize("name", ALL)
def test_against_independent_brute_force(make, name):
    """Third guard, sharing no code or formulation with the oracle (tests/bruteforce.py: world-frame 3-vectors, point
    Jacobians, kinetic-energy mass matrix, projected Newton-Euler, dense Cholesky solve, numpy 80-bit floats):
    M == crba (unittest/crba.cpp:106-114 asks the same of rnea columns), tau == rnea, solve(M, tau - b) == aba
    (unittest/aba.cpp:143-154), at the reference's 1e-12."""
    from bruteforce import BruteForce
    m, o = make(name)
    bf = BruteForce(m)
    q, v, a = random_inputs(m, 3, 12)
    for i in range(3):
        qi, vi, ai = q[:, i], v[:, i], a[:, i]
        M = bf.mass_matrix(qi).astype(np.float64)
        assert is_approx(sym_from_upper(mat(o.crba(qi, world=True)[:, 0], m.nv)), M, 1e-12), (name, i)
        assert is_approx(sym_from_upper(mat(o.crba(qi, world=False)[:, 0], m.nv)), M, 1e-12), (name, i)
        tau = bf.inverse_dynamics(qi, vi, ai).astype(np.float64)
        assert is_approx(o.rnea(qi, vi, ai)[:, 0], tau, 1e-12), (name, i)
        ddq = bf.forward_dynamics(qi, vi, ai)[0].astype(np.float64)
        assert is_approx(o.aba(qi, vi, ai)[:, 0], ddq, 1e-12), (name, i)
        # the analytical derivatives against central differences of the brute force (no oracle on the right-hand side)
    qi, vi, ai = q[:, 0], v[:, 0], a[:, 0]
    if m.nv <= 12:
        h = 1e-6
        dv = mat(o.rnea_derivatives(qi, vi, ai)[1][:, 0], m.nv)
        fd = np.zeros((m.nv, m.nv))
        for k in range(m.nv):
            e = np.zeros(m.nv)
            e[k] = h
            fd[:, k] = ((bf.inverse_dynamics(qi, vi + e, ai) - bf.inverse_dynamics(qi, vi - e, ai)) / (2 * h)).astype(np.float64)
        assert is_approx(dv, fd, 1e-8), name


def test_parallel_equals_serial(make):
    """unittest/parallel-rnea.cpp:21-55 / parallel-aba.cpp:21-55: bit-exact."""
    m, o = make("humanoid_random")
    q, v, a = random_inputs(m, 128, 9)
    t1, tn = o.rnea(q, v, a, nthreads=1), o.rnea(q, v, a, nthreads=4)
    assert np.array_equal(t1, tn)
    a1, an = o.aba(q, v, t1, nthreads=1), o.aba(q, v, t1, nthreads=3)
    assert np.array_equal(a1, an)


def test_pendulum_closed_form(oracle_cls):
    """Analytic guard: a point mass m at distance l on a revolute-Y joint: tau = m l^2 a + m g l sin(q)."""
    from pinocchio_b200 import model as M
    mdl = M.Model()
    j = mdl.addJoint(0, M.JOINT_RY, M.SE3.Identity(), "pend", [-3.0], [3.0])
    mass, l = 2.0, 0.7
    mdl.appendBodyToJoint(j, M.Inertia(mass, [0.0, 0.0, -l], np.zeros((3, 3))))
    o = oracle_cls(mdl)
    for qv, av in ((0.3, 0.0), (-1.1, 2.0), (2.0, -0.5)):
        tau = o.rnea(np.array([qv]), np.array([0.4]), np.array([av]))[0, 0]
        assert abs(tau - (mass * l * l * av + mass * 9.81 * l * np.sin(qv))) < 1e-12
        Mq = o.crba(np.array([qv]))[0, 0]
        assert abs(Mq - mass * l * l) < 1e-13


def test_flop_counts_are_positive_and_ordered(make):
    m, o = make("humanoid_random")
    q, v, a = random_inputs(m, 1, 10)
    c = {k: o.count_flops(k, q[:, 0], v[:, 0], a[:, 0]) for k in o.ALGOS}
    assert c["rnea"]["flops"] < c["aba"]["flops"] < c["rnea_derivatives"]["flops"] < c["aba_derivatives"]["flops"]
    assert c["rnea"]["sincos"] == 26


# ---- integrate (the step after ABA) ----------------------------------------------------------
@pytest.mark.parametrize("name", ALL)
def test_integrate_identities(make, name):
    """unittest/joint-configurations.cpp:46-62 (integration_test): integrate(q, 0) == q; plus exp(v) exp(-v) == 1 on every
    Lie group, unit-norm quaternions / unit complex after the step, and agreement with the independent host-side
    restatement pinocchio_b200.joint_configuration.integrate and with the long double instantiation."""
    import pinocchio_b200 as pb
    m, o = make(name)
    q, v, _ = random_inputs(m, 8, 21)
    assert np.array_equal(o.integrate(q, 0.0 * v), q) or np.abs(o.integrate(q, 0.0 * v) - q).max() < 2e-16
    q1 = o.integrate(q, 0.7 * v)
    assert np.abs(o.integrate(q1, -0.7 * v) - q).max() < 1e-14
    assert np.abs(o.integrate(q, 0.7 * v, long_double=True) - q1).max() < 1e-14
    for b in range(q.shape[1]):
        assert np.abs(pb.integrate(m, q[:, b], 0.7 * v[:, b]) - q1[:, b]).max() < 1e-13
    for j in range(1, m.njoints):
        iq, t = m.idx_qs[j], m.joint_types[j]
        if t == pb.JOINT_FREEFLYER:
            assert np.abs(np.linalg.norm(q1[iq + 3:iq + 7], axis=0) - 1).max() < 1e-14
        elif t == pb.JOINT_SPHERICAL:
            assert np.abs(np.linalg.norm(q1[iq:iq + 4], axis=0) - 1).max() < 1e-14
        elif t == pb.JOINT_PLANAR:
            assert np.abs(np.linalg.norm(q1[iq + 2:iq + 4], axis=0) - 1).max() < 1e-14


def test_integrate_known_answers(oracle_cls):
    """Closed forms: a free-flyer at the identity moved by the twist (v, w) = ((1, 0, 0), (0, 0, th)) ends on the arc
    (sin th / th, (1 - cos th) / th, 0) with the quaternion (0, 0, sin th/2, cos th/2) (exp6, explog-quaternion.hpp:92-136);
    the small-angle branch (Taylor expansion below epsilon^(1/4)) agrees with the closed form to first order."""
    import pinocchio_b200 as pb
    from pinocchio_b200 import model as M
    m = M.Model()
    f = m.addJoint(0, M.JOINT_FREEFLYER, M.SE3.Identity(), "ff", np.full(7, -1.0), np.full(7, 1.0))
    m.appendBodyToJoint(f, M._Rng(3).inertia(), M.SE3.Identity())
    o = oracle_cls(m)
    q0 = np.array([0, 0, 0, 0, 0, 0, 1.0]).reshape(7, 1)
    for th in (0.9, 1e-3, 1e-7):
        v = np.array([1.0, 0, 0, 0, 0, th]).reshape(6, 1)
        q1 = o.integrate(q0, v)[:, 0]
        ref = np.array([np.sin(th) / th, 2 * np.sin(th / 2) ** 2 / th, 0, 0, 0, np.sin(th / 2), np.cos(th / 2)])
        assert np.abs(q1 - ref).max() < 1e-12, (th, q1, ref)


def test_unaligned_joints_reduce_to_the_aligned_ones(oracle_cls):
    """unittest/joint-revolute.cpp / joint-prismatic.cpp "vsRX / vsPX" pattern: JointModelRevoluteUnaligned(e_k) gives the
    results of JointModelR{X,Y,Z}, PrismaticUnaligned(e_k) those of P{X,Y,Z} — pins the oracle's unaligned branch against
    its axis-aligned one (joint-revolute-unaligned.hpp:668-672 vs joint-revolute.hpp:791-820)."""
    from pinocchio_b200 import model as M

    def chain(unaligned):
        rng = M._Rng(21)
        m = M.Model()
        parent = 0
        for k in range(6):
            ax = np.eye(3)[k % 3]
            rev = k < 3
            if unaligned:
                jt = M.JOINT_REVOLUTE_UNALIGNED if rev else M.JOINT_PRISMATIC_UNALIGNED
                parent = m.addJoint(parent, jt, rng.se3(), f"j{k}", [-1.0], [1.0], axis=ax)
            else:
                parent = m.addJoint(parent, (M.JOINT_RX if rev else M.JOINT_PX) + k % 3, rng.se3(), f"j{k}", [-1.0], [1.0])
            m.appendBodyToJoint(parent, rng.inertia(), M.SE3.Identity())
        return m
    ma, mu = chain(False), chain(True)
    oa, ou = oracle_cls(ma), oracle_cls(mu)
    q, v, a = random_inputs(ma, 5, 3)
    tau = oa.rnea(q, v, a)
    assert np.allclose(ou.rnea(q, v, a), tau, rtol=0, atol=1e-13 * np.abs(tau).max())
    assert np.allclose(ou.aba(q, v, tau), oa.aba(q, v, tau), rtol=0, atol=1e-11 * np.abs(a).max())
    assert np.allclose(ou.crba(q, world=True), oa.crba(q, world=True), rtol=0, atol=1e-13 * np.abs(tau).max())
    for x, y in zip(ou.rnea_derivatives(q, v, a), oa.rnea_derivatives(q, v, a)):
        assert np.allclose(x, y, rtol=0, atol=1e-12 * max(1.0, np.abs(y).max()))
