"""Independent brute-force rigid-body dynamics in numpy extended precision (TEST INFRASTRUCTURE).

Shares NO code and no formulation with oracle/rbd_oracle.hpp or the CUDA kernels: no spatial (6-D) algebra, no recursion
over the tree, no joint-space sparsity.  Everything is world-frame 3-vectors, textbook style:

* forward kinematics: one rotation matrix / origin per joint, composed down the tree;
* every tangent coordinate k is an *axis* a_k fixed in the child body of its joint, through the joint origin o_k,
  either translational (v += a_k qdot_k) or rotational (omega += a_k qdot_k, v(x) += a_k x (x - o_k) qdot_k);
* M(q) from the kinetic energy:  M_kl = sum_b  m_b Jv_k(c_b).Jv_l(c_b) + Jw_k . I_b Jw_l   (point Jacobians at the CoM);
* tau(q, qd, qdd) by projected Newton-Euler (Kane): F_b = m_b (a(c_b) - g), N_b = I_b alpha_b + omega_b x I_b omega_b,
  tau_k = sum_b  Jv_k(c_b).F_b + Jw_k.N_b, with the accelerations obtained by differentiating the velocity sums term by term
  (d a_k / dt = omega_body(k) x a_k);
* qdd = solve(M, tau - tau(q, qd, 0)).

np.longdouble is the x87 80-bit type on the x86 hosts this runs on (64-bit mantissa), so the result is also a
higher-precision check.  O(n^2) python loops: meant for a handful of configurations.
The reference identities it stands in for: unittest/crba.cpp:106-114 (M column by column), unittest/aba.cpp:143-154.
"""
import numpy as np

LD = np.longdouble

# joint type tags as in include/pinocchio_b200.h (duplicated on purpose: no import from the package's algorithms)
RX, RY, RZ, PX, PY, PZ, FF, SPH, PLANAR, RU, PU, RUBX, RUBY, RUBZ, RUBU = range(15)


def _quat_R(x, y, z, w):
    """Rotation matrix of the unit quaternion (x, y, z, w), from R = (w^2 - v.v) 1 + 2 v v^T + 2 w [v]x."""
    v = np.array([x, y, z], dtype=LD)
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=LD)
    return (w * w - v @ v) * np.eye(3, dtype=LD) + 2 * np.outer(v, v) + 2 * w * K


def _axis_R(axis, angle, cs=None):
    """Rodrigues; `cs` = (cos, sin) given directly (unbounded joints)."""
    a = np.asarray(axis, dtype=LD)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dtype=LD)
    c, s_ = (np.cos(angle), np.sin(angle)) if cs is None else cs
    return np.eye(3, dtype=LD) + s_ * K + (1 - c) * (K @ K)


class BruteForce:
    def __init__(self, model):
        self.m = model
        m = model
        self.nv, self.nj = m.nv, m.njoints
        self.anc = [[] for _ in range(m.njoints)]  # ancestors-or-self, root first
        for j in range(1, m.njoints):
            self.anc[j] = self.anc[m.parents[j]] + [j]

    def _kinematics(self, q):
        m = self.m
        q = np.asarray(q, dtype=LD)
        R = [np.eye(3, dtype=LD) for _ in range(self.nj)]
        p = [np.zeros(3, dtype=LD) for _ in range(self.nj)]
        dofs = []  # (joint, kind 'l'/'a', axis in world, origin in world)
        E = np.eye(3, dtype=LD)
        for j in range(1, self.nj):
            t, par = m.joint_types[j], m.parents[j]
            Rp = np.asarray(m.jointPlacements[j].R, dtype=LD)
            pp = np.asarray(m.jointPlacements[j].p, dtype=LD)
            qj = q[m.idx_qs[j]:m.idx_qs[j] + m.nqs[j]]
            if t in (RX, RY, RZ):
                Rj, pj, ax = _axis_R(E[t - RX], qj[0]), np.zeros(3, dtype=LD), [("a", E[t - RX])]
            elif t in (PX, PY, PZ):
                Rj, pj, ax = E, qj[0] * E[t - PX], [("l", E[t - PX])]
            elif t == RU:
                a = np.asarray(m.axes[j], dtype=LD)
                Rj, pj, ax = _axis_R(a, qj[0]), np.zeros(3, dtype=LD), [("a", a)]
            elif t == PU:
                a = np.asarray(m.axes[j], dtype=LD)
                Rj, pj, ax = E, qj[0] * a, [("l", a)]
            elif t in (RUBX, RUBY, RUBZ):
                Rj, pj, ax = _axis_R(E[t - RUBX], None, (qj[0], qj[1])), np.zeros(3, dtype=LD), [("a", E[t - RUBX])]
            elif t == RUBU:
                a = np.asarray(m.axes[j], dtype=LD)
                Rj, pj, ax = _axis_R(a, None, (qj[0], qj[1])), np.zeros(3, dtype=LD), [("a", a)]
            elif t == FF:
                Rj, pj = _quat_R(qj[3], qj[4], qj[5], qj[6]), qj[:3].copy()
                ax = [("l", E[0]), ("l", E[1]), ("l", E[2]), ("a", E[0]), ("a", E[1]), ("a", E[2])]
            elif t == SPH:
                Rj, pj, ax = _quat_R(qj[0], qj[1], qj[2], qj[3]), np.zeros(3, dtype=LD), [("a", E[0]), ("a", E[1]), ("a", E[2])]
            elif t == PLANAR:
                c, s = qj[2], qj[3]
                Rj = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=LD)
                pj, ax = np.array([qj[0], qj[1], 0], dtype=LD), [("l", E[0]), ("l", E[1]), ("a", E[2])]
            else:
                raise ValueError(t)
            R[j] = R[par] @ Rp @ Rj
            p[j] = p[par] + R[par] @ (pp + Rp @ pj)
            # tangent axes are expressed in the CHILD frame (body-frame velocities) and pass through the child origin
            for kind, a in ax:
                dofs.append((j, kind, R[j] @ a, p[j]))
        assert len(dofs) == self.nv
        return R, p, dofs

    def _bodies(self, R, p):
        m = self.m
        out = []
        for j in range(1, self.nj):
            Y = m.inertias[j]
            s = np.asarray(Y.sym, dtype=LD)  # xx, xy, yy, xz, yz, zz
            Ic = np.array([[s[0], s[1], s[3]], [s[1], s[2], s[4]], [s[3], s[4], s[5]]], dtype=LD)
            out.append((j, LD(Y.mass), p[j] + R[j] @ np.asarray(Y.lever, dtype=LD), R[j] @ Ic @ R[j].T))
        return out

    def _on_path(self, k_joint, body_joint):
        return k_joint in self.anc[body_joint]

    def mass_matrix(self, q):
        R, p, dofs = self._kinematics(q)
        nv = self.nv
        M = np.zeros((nv, nv), dtype=LD)
        for (b, mass, c, I) in self._bodies(R, p):
            Jv = np.zeros((3, nv), dtype=LD)
            Jw = np.zeros((3, nv), dtype=LD)
            for k, (j, kind, a, o) in enumerate(dofs):
                if not self._on_path(j, b):
                    continue
                if kind == "l":
                    Jv[:, k] = a
                else:
                    Jv[:, k] = np.cross(a, c - o)
                    Jw[:, k] = a
            M += mass * (Jv.T @ Jv) + Jw.T @ I @ Jw
        M += np.diag(np.asarray(self.m.armature, dtype=LD))
        return M

    def inverse_dynamics(self, q, qd, qdd):
        R, p, dofs = self._kinematics(q)
        qd, qdd = np.asarray(qd, dtype=LD), np.asarray(qdd, dtype=LD)
        nv = self.nv
        g = np.asarray(self.m.gravity, dtype=LD)
        # angular velocity of every joint's body
        omega = [np.zeros(3, dtype=LD) for _ in range(self.nj)]
        for j in range(1, self.nj):
            for k, (jk, kind, a, o) in enumerate(dofs):
                if kind == "a" and self._on_path(jk, j):
                    omega[j] = omega[j] + a * qd[k]

        def vel(x, body):  # velocity of the point x fixed on `body`
            v = np.zeros(3, dtype=LD)
            for k, (jk, kind, a, o) in enumerate(dofs):
                if self._on_path(jk, body):
                    v = v + (a * qd[k] if kind == "l" else np.cross(a, x - o) * qd[k])
            return v

        vo = [vel(o, jk) for (jk, kind, a, o) in dofs]  # velocity of each axis origin, as a point of its own body

        def acc(x, body):
            vx = vel(x, body)
            ac = np.zeros(3, dtype=LD)
            for k, (jk, kind, a, o) in enumerate(dofs):
                if not self._on_path(jk, body):
                    continue
                adot = np.cross(omega[jk], a)
                if kind == "l":
                    ac = ac + adot * qd[k] + a * qdd[k]
                else:
                    ac = ac + (np.cross(adot, x - o) + np.cross(a, vx - vo[k])) * qd[k] + np.cross(a, x - o) * qdd[k]
            return ac

        tau = np.zeros(nv, dtype=LD)
        for (b, mass, c, I) in self._bodies(R, p):
            alpha = np.zeros(3, dtype=LD)
            for k, (jk, kind, a, o) in enumerate(dofs):
                if kind == "a" and self._on_path(jk, b):
                    alpha = alpha + np.cross(omega[jk], a) * qd[k] + a * qdd[k]
            F = mass * (acc(c, b) - g)
            N = I @ alpha + np.cross(omega[b], I @ omega[b])
            for k, (jk, kind, a, o) in enumerate(dofs):
                if not self._on_path(jk, b):
                    continue
                tau[k] += a @ F if kind == "l" else a @ N + np.cross(a, c - o) @ F
        return tau + np.asarray(self.m.armature, dtype=LD) * qdd

    def forward_dynamics(self, q, qd, tau):
        M = self.mass_matrix(q)
        b = self.inverse_dynamics(q, qd, np.zeros(self.nv))
        return _solve_spd(M, np.asarray(tau, dtype=LD) - b), M, b


def _solve_spd(A, rhs):
    """Cholesky solve in extended precision (numpy.linalg has no longdouble kernels)."""
    n = A.shape[0]
    L = np.zeros_like(A)
    for i in range(n):
        for j in range(i + 1):
            s = A[i, j] - L[i, :j] @ L[j, :j]
            L[i, j] = np.sqrt(s) if i == j else s / L[j, j]
    y = np.zeros(n, dtype=LD)
    for i in range(n):
        y[i] = (rhs[i] - L[i, :i] @ y[:i]) / L[i, i]
    x = np.zeros(n, dtype=LD)
    for i in reversed(range(n)):
        x[i] = (y[i] - L[i + 1:, i] @ x[i + 1:]) / L[i, i]
    return x
