"""Host-side input generation and manifold integration (test / benchmark utilities).

Restates, for the five supported joint families:

* ``randomConfiguration(model, lo, hi)`` — include/pinocchio/algorithm/joint-configuration.hxx:158-186;
  vector-space coordinates ``lo + (hi-lo)*rand()/RAND_MAX`` (multibody/liegroup/vector-space.hpp:293-313),
  quaternions uniform on S^3 whatever the bounds (multibody/liegroup/special-orthogonal.hpp:674-688,
  math/quaternion.hpp:115-137), free-flyer = R^3 x SO(3) (special-euclidean.hpp:910-916).
* ``integrate(model, q, v)`` — needed by the finite-difference derivative tests
  (unittest/rnea-derivatives.cpp:158-171): vector space ``q+v``; SO(3) / SE(3) through the
  exponential map (special-orthogonal.hpp, special-euclidean.hpp:660-698, spatial/explog.hpp:347-405).
* ``neutral(model)``.

Two generators: ``LibcRand`` replays libc ``srand/rand`` in the reference's consumption order
(small batches, fidelity); ``batched_random_configuration`` is the vectorised Philox generator
used for >= 1M-column batches (seed recorded by the caller).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
from typing import Optional

import numpy as np

from .model import (JOINT_FREEFLYER, JOINT_PLANAR, JOINT_PZ, JOINT_REVOLUTE_UNALIGNED, JOINT_SPHERICAL, Model, joint_is_unbounded)

RAND_MAX = 2147483647


class LibcRand:
    """glibc ``srand``/``rand`` — the generator behind the reference's Random() calls."""

    def __init__(self, seed: int = 0):
        self._libc = ctypes.CDLL(ctypes.util.find_library("c") or "libc.so.6")
        self._libc.rand.restype = ctypes.c_int
        self._libc.srand(ctypes.c_uint(seed))

    def rand(self) -> int:
        return int(self._libc.rand())

    def unit(self) -> float:
        return self.rand() / RAND_MAX

    def eigen_random(self, n: int) -> np.ndarray:
        """Eigen 3.4 ``DenseBase::Random`` for double: uniform in [-1, 1] from one rand() each."""
        return np.array([2.0 * self.rand() / RAND_MAX - 1.0 for _ in range(n)])


def _uniform_quaternion(u1, u2, u3):
    """math/quaternion.hpp:115-137; returns (x, y, z, w)."""
    m1, m2 = np.sqrt(1.0 - u1), np.sqrt(u1)
    s2, c2 = np.sin(2.0 * math.pi * u2), np.cos(2.0 * math.pi * u2)
    s3, c3 = np.sin(2.0 * math.pi * u3), np.cos(2.0 * math.pi * u3)
    return m1 * c2, m2 * s3, m2 * c3, m1 * s2


def randomConfiguration(model: Model, lo=None, hi=None, rng: Optional[LibcRand] = None) -> np.ndarray:
    """One configuration, joints visited in index order (joint-configuration.hxx:177-185)."""
    rng = LibcRand(0) if rng is None else rng
    lo = model.lowerPositionLimit if lo is None else np.broadcast_to(np.asarray(lo, dtype=np.float64), (model.nq,))
    hi = model.upperPositionLimit if hi is None else np.broadcast_to(np.asarray(hi, dtype=np.float64), (model.nq,))
    q = np.zeros(model.nq)

    def vec(i0, n):
        for k in range(i0, i0 + n):
            if not (np.isfinite(lo[k]) and np.isfinite(hi[k])):
                raise ValueError(f"non bounded limit. Cannot uniformly sample joint at rank {k}")
            q[k] = lo[k] + ((hi[k] - lo[k]) * rng.rand()) / RAND_MAX

    for j in range(1, model.njoints):
        t, iq = model.joint_types[j], model.idx_qs[j]
        if joint_is_unbounded(t):
            # SpecialOrthogonalOperationTpl<2>::random_impl (special-orthogonal.hpp): angle in [-pi, pi] -> (cos, sin)
            ang = -math.pi + 2.0 * math.pi * rng.unit()
            q[iq], q[iq + 1] = math.cos(ang), math.sin(ang)
        elif t <= JOINT_PZ or t >= JOINT_REVOLUTE_UNALIGNED:
            vec(iq, 1)
        elif t == JOINT_FREEFLYER:
            vec(iq, 3)
            q[iq + 3:iq + 7] = _uniform_quaternion(rng.unit(), rng.unit(), rng.unit())
        elif t == JOINT_SPHERICAL:
            q[iq:iq + 4] = _uniform_quaternion(rng.unit(), rng.unit(), rng.unit())
        elif t == JOINT_PLANAR:
            # R^2 x SO(2): special-orthogonal.hpp (SO(2) random: angle in [-pi, pi] -> (cos, sin))
            vec(iq, 2)
            ang = -math.pi + 2.0 * math.pi * rng.unit()
            q[iq + 2], q[iq + 3] = math.cos(ang), math.sin(ang)
    return q


def batched_random_configuration(model: Model, batch: int, seed: int, lo: float = -1.0, hi: float = 1.0,
                                 dtype=np.float64) -> np.ndarray:
    """(nq x batch) Fortran-ordered configurations from a counter-based Philox stream.

    Same distributions as ``randomConfiguration(model, -1, +1)`` in benchmark/timings-parallel.cpp:48,57.
    """
    g = np.random.Generator(np.random.Philox(seed))
    q = np.empty((batch, model.nq), dtype=np.float64)
    for j in range(1, model.njoints):
        t, iq = model.joint_types[j], model.idx_qs[j]
        if joint_is_unbounded(t):
            ang = g.uniform(-math.pi, math.pi, batch)
            q[:, iq], q[:, iq + 1] = np.cos(ang), np.sin(ang)
        elif t <= JOINT_PZ or t >= JOINT_REVOLUTE_UNALIGNED:
            q[:, iq] = g.uniform(lo, hi, batch)
        elif t == JOINT_FREEFLYER:
            q[:, iq:iq + 3] = g.uniform(lo, hi, (batch, 3))
            u = g.random((batch, 3))
            x, y, z, w = _uniform_quaternion(u[:, 0], u[:, 1], u[:, 2])
            q[:, iq + 3], q[:, iq + 4], q[:, iq + 5], q[:, iq + 6] = x, y, z, w
        elif t == JOINT_SPHERICAL:
            u = g.random((batch, 3))
            x, y, z, w = _uniform_quaternion(u[:, 0], u[:, 1], u[:, 2])
            q[:, iq], q[:, iq + 1], q[:, iq + 2], q[:, iq + 3] = x, y, z, w
        else:
            q[:, iq:iq + 2] = g.uniform(lo, hi, (batch, 2))
            ang = g.uniform(-math.pi, math.pi, batch)
            q[:, iq + 2], q[:, iq + 3] = np.cos(ang), np.sin(ang)
    return np.asfortranarray(q.T.astype(dtype))


def batched_random_tangent(model: Model, batch: int, seed: int, dtype=np.float64) -> np.ndarray:
    """(nv x batch) U(-1, 1), as ``VectorXd::Random`` in benchmark/timings-parallel.cpp:58-60."""
    g = np.random.Generator(np.random.Philox(seed))
    return np.asfortranarray(g.uniform(-1.0, 1.0, (batch, model.nv)).T.astype(dtype))


def neutral(model: Model) -> np.ndarray:
    q = np.zeros(model.nq)
    for j in range(1, model.njoints):
        t, iq = model.joint_types[j], model.idx_qs[j]
        if t == JOINT_FREEFLYER:
            q[iq + 6] = 1.0
        elif t == JOINT_SPHERICAL:
            q[iq + 3] = 1.0
        elif t == JOINT_PLANAR:
            q[iq + 2] = 1.0
        elif joint_is_unbounded(t):
            q[iq] = 1.0  # (cos, sin) of a zero angle
    return q


# --------------------------------------------------------------------------------------------
# integrate
# --------------------------------------------------------------------------------------------
def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


def _quat_rotate(q, v):
    x, y, z, w = q
    u = np.array([x, y, z])
    return v + 2.0 * np.cross(u, np.cross(u, v) + w * v)


def _exp3_quat(w):
    """quaternion::exp3 — unit quaternion (x, y, z, w) of the rotation vector ``w``."""
    t2 = float(w @ w)
    t = math.sqrt(t2)
    if t < 1e-8:
        k = 0.5 - t2 / 48.0
        return np.array([k * w[0], k * w[1], k * w[2], 1.0 - t2 / 8.0])
    s = math.sin(0.5 * t) / t
    return np.array([s * w[0], s * w[1], s * w[2], math.cos(0.5 * t)])


def _exp6(v):
    """SE(3) exponential of a body twist (linear, angular) -> (translation, quaternion)."""
    lin, w = v[:3], v[3:]
    t2 = float(w @ w)
    t = math.sqrt(t2)
    if t < 1e-8:
        alpha_wxv = 0.5 - t2 / 24.0
        alpha_v = 1.0 - t2 / 6.0
        alpha_w = 1.0 / 6.0 - t2 / 120.0
    else:
        st, ct = math.sin(t), math.cos(t)
        alpha_wxv = (1.0 - ct) / t2
        alpha_v = st / t
        alpha_w = (1.0 - alpha_v) / t2
    trans = alpha_v * lin + alpha_w * float(w @ lin) * w + alpha_wxv * np.cross(w, lin)
    return trans, _exp3_quat(w)


def _first_order_normalize(q):  # math/quaternion.hpp:90-99
    n2 = float(q @ q)
    return q * ((3.0 - n2) / 2.0)


def integrate(model: Model, q: np.ndarray, v: np.ndarray) -> np.ndarray:
    """q (+) v on the configuration manifold, joint by joint."""
    out = np.array(q, dtype=np.float64, copy=True)
    for j in range(1, model.njoints):
        t, iq, iv = model.joint_types[j], model.idx_qs[j], model.idx_vs[j]
        if joint_is_unbounded(t):  # SO(2), special-orthogonal.hpp:164-185
            ca, sa = q[iq], q[iq + 1]
            co, so = math.cos(v[iv]), math.sin(v[iv])
            c1, s1 = co * ca - so * sa, so * ca + co * sa
            n = (3.0 - (c1 * c1 + s1 * s1)) / 2.0
            out[iq], out[iq + 1] = c1 * n, s1 * n
        elif t <= JOINT_PZ or t >= JOINT_REVOLUTE_UNALIGNED:
            out[iq] = q[iq] + v[iv]
        elif t == JOINT_FREEFLYER:  # special-euclidean.hpp:660-698
            quat = q[iq + 3:iq + 7]
            trans, dq = _exp6(v[iv:iv + 6])
            out[iq:iq + 3] = _quat_rotate(quat, trans) + q[iq:iq + 3]
            res = _quat_mul(quat, dq)
            if float(res @ quat) < 0.0:
                res = -res
            out[iq + 3:iq + 7] = _first_order_normalize(res)
        elif t == JOINT_SPHERICAL:
            # SpecialOrthogonalOperationTpl<3>::integrate_impl (special-orthogonal.hpp:467-481): no sign fix-up (only SE(3) has one)
            quat = q[iq:iq + 4]
            res = _quat_mul(quat, _exp3_quat(v[iv:iv + 3]))
            out[iq:iq + 4] = _first_order_normalize(res)
        else:  # planar: SE(2), q = (x, y, cos, sin), v = (vx, vy, wz) body frame
            c0, s0 = q[iq + 2], q[iq + 3]
            vx, vy, w = v[iv:iv + 3]
            if abs(w) > 1e-14:
                sw, cw = math.sin(w), math.cos(w)
                tx = (sw * vx - (1.0 - cw) * vy) / w
                ty = ((1.0 - cw) * vx + sw * vy) / w
            else:
                sw, cw = w, 1.0
                tx, ty = vx, vy
            out[iq] = q[iq] + c0 * tx - s0 * ty
            out[iq + 1] = q[iq + 1] + s0 * tx + c0 * ty
            # out.tail<2>() = R0 * R.col(0) (special-euclidean.hpp:304-305): (cos, sin) is not re-normalised
            out[iq + 2], out[iq + 3] = c0 * cw - s0 * sw, s0 * cw + c0 * sw
    return out
