"""ctypes wrapper around oracle/liboracle.so (TEST INFRASTRUCTURE ONLY).

Parity status: "parity unpinned" by golden vectors (the reference has none for this path);
pinned by the reference's identity tests, see tests/test_oracle_identities.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _FlatModel(ctypes.Structure):
    _fields_ = [
        ("njoints", ctypes.c_int32), ("nq", ctypes.c_int32), ("nv", ctypes.c_int32),
        ("parents", ctypes.POINTER(ctypes.c_int32)), ("joint_type", ctypes.POINTER(ctypes.c_int32)),
        ("idx_q", ctypes.POINTER(ctypes.c_int32)), ("idx_v", ctypes.POINTER(ctypes.c_int32)),
        ("placement", ctypes.POINTER(ctypes.c_double)), ("inertia", ctypes.POINTER(ctypes.c_double)),
        ("armature", ctypes.POINTER(ctypes.c_double)), ("gravity", ctypes.c_double * 3),
        ("axis", ctypes.POINTER(ctypes.c_double)),
    ]


def build_oracle(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc -O3 -fopenmp)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "rbd_oracle.hpp")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_oracle())
        L.oracle_model_create.restype = ctypes.c_void_p
        L.oracle_model_create.argtypes = [ctypes.POINTER(_FlatModel)]
        L.oracle_model_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _nt(nthreads, B):
    """0 = every host thread for batches worth forking a team for, one thread otherwise."""
    if nthreads and nthreads > 0:
        return int(nthreads)
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return n if B >= 512 else 1


def _cols(a, rows):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(rows, 1)
    assert a.shape[0] == rows, (a.shape, rows)
    return np.asfortranarray(a)


class Oracle:
    """CPU restatement of rnea / aba(WORLD) / crba / computeRNEADerivatives / computeABADerivatives.

    All batched arguments are (rows x B) arrays whose columns are configurations, as in
    rneaInParallel (include/pinocchio/algorithm/parallel/rnea.hpp:38).  `long_double`: 0 / False = double,
    1 / True = x87 80-bit, 2 = IEEE binary128 (__float128); `nthreads` 0 = all host threads for B >= 512.
    """

    def __init__(self, model):
        f = model.flat() if hasattr(model, "flat") else model
        self._keep = f
        self.nq, self.nv, self.njoints = int(f["nq"]), int(f["nv"]), int(f["njoints"])
        fm = _FlatModel()
        fm.njoints, fm.nq, fm.nv = self.njoints, self.nq, self.nv
        ip = lambda k: f[k].ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        dp = lambda k: f[k].ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        fm.parents, fm.joint_type, fm.idx_q, fm.idx_v = ip("parents"), ip("joint_type"), ip("idx_q"), ip("idx_v")
        fm.placement, fm.inertia, fm.armature = dp("placement"), dp("inertia"), dp("armature")
        if "axis" not in f:
            f = dict(f)
            f["axis"] = np.zeros(3 * self.njoints)
            self._keep = f
        f["axis"] = np.ascontiguousarray(f["axis"], dtype=np.float64)
        fm.axis = dp("axis")
        for k in range(3):
            fm.gravity[k] = float(f["gravity"][k])
        self._h = ctypes.c_void_p(_lib().oracle_model_create(ctypes.byref(fm)))

    def __del__(self):
        try:
            if self._h:
                _lib().oracle_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def max_threads() -> int:
        return int(_lib().oracle_max_threads())

    def rnea(self, q, v, a, nthreads=0, long_double=False, out=None):
        q, v, a = _cols(q, self.nq), _cols(v, self.nv), _cols(a, self.nv)
        B = q.shape[1]
        tau = np.empty((self.nv, B), order="F") if out is None else out
        _lib().oracle_rnea(self._h, _p(q), _p(v), _p(a), _p(tau), ctypes.c_int64(B), _nt(nthreads, B), int(long_double))
        return tau

    def aba(self, q, v, tau, nthreads=0, long_double=False, out=None):
        q, v, tau = _cols(q, self.nq), _cols(v, self.nv), _cols(tau, self.nv)
        B = q.shape[1]
        a = np.empty((self.nv, B), order="F") if out is None else out
        _lib().oracle_aba(self._h, _p(q), _p(v), _p(tau), _p(a), ctypes.c_int64(B), _nt(nthreads, B), int(long_double))
        return a

    def crba(self, q, nthreads=0, long_double=False, world=False, out=None):
        """(nv*nv x B); each column a col-major nv x nv matrix, upper triangle + zeros.  `out`: caller-owned F-order block
        (the timed CPU arm of bench.py reuses one, as a caller of the reference would reuse its Data)."""
        q = _cols(q, self.nq)
        B = q.shape[1]
        M = np.empty((self.nv * self.nv, B), order="F") if out is None else out
        _lib().oracle_crba(self._h, _p(q), _p(M), ctypes.c_int64(B), _nt(nthreads, B), int(long_double), int(world))
        return M

    def rnea_derivatives(self, q, v, a, nthreads=0, long_double=False, out=None):
        q, v, a = _cols(q, self.nq), _cols(v, self.nv), _cols(a, self.nv)
        B = q.shape[1]
        nn = self.nv * self.nv
        dq, dv, da, tau = out if out is not None else (*(np.empty((nn, B), order="F") for _ in range(3)), np.empty((self.nv, B), order="F"))
        _lib().oracle_rnea_derivatives(self._h, _p(q), _p(v), _p(a), _p(dq), _p(dv), _p(da), _p(tau),
                                       ctypes.c_int64(B), _nt(nthreads, B), int(long_double))
        return dq, dv, da, tau

    def aba_derivatives(self, q, v, tau, nthreads=0, long_double=False, out=None):
        q, v, tau = _cols(q, self.nq), _cols(v, self.nv), _cols(tau, self.nv)
        B = q.shape[1]
        nn = self.nv * self.nv
        dq, dv, dtau, ddq = out if out is not None else (*(np.empty((nn, B), order="F") for _ in range(3)), np.empty((self.nv, B), order="F"))
        _lib().oracle_aba_derivatives(self._h, _p(q), _p(v), _p(tau), _p(dq), _p(dv), _p(dtau), _p(ddq),
                                      ctypes.c_int64(B), _nt(nthreads, B), int(long_double))
        return dq, dv, dtau, ddq

    # ---- the callers' other needs ----
    def nle(self, q, v, nthreads=0):
        """nonLinearEffects (algorithm/rnea.hxx:227-343): the same two sweeps as rnea with the S*a term dropped, i.e.
        rnea(q, v, 0) — the identity the reference asserts in unittest/rnea.cpp:201-206."""
        q, v = _cols(q, self.nq), _cols(v, self.nv)
        return self.rnea(q, v, np.zeros_like(v), nthreads=nthreads)

    def gravity(self, q, nthreads=0):
        """computeGeneralizedGravity (algorithm/rnea.hxx:346-452) == rnea(q, 0, 0), unittest/rnea.cpp:225-228."""
        q = _cols(q, self.nq)
        z = np.zeros((self.nv, q.shape[1]), order="F")
        return self.rnea(q, z, z, nthreads=nthreads)

    def minverse(self, q, nthreads=0):
        """computeMinverse (algorithm/aba.hxx:613-902): upper triangle of M^-1, strictly-lower part zero (a fresh
        data.Minv).  Taken from the Minv recursion of abaDerivatives, which the reference asserts equal to
        computeMinverse (unittest/aba-derivatives.cpp:96-100); it does not depend on v or tau."""
        q = _cols(q, self.nq)
        z = np.zeros((self.nv, q.shape[1]), order="F")
        Minv = self.aba_derivatives(q, z, z, nthreads=nthreads)[2]
        nv = self.nv
        mask = np.triu(np.ones((nv, nv), dtype=bool)).reshape(-1, order="F")
        return np.asfortranarray(Minv * mask[:, None])

    def integrate(self, q, v, long_double=False):
        """integrate(model, q, v) (algorithm/joint-configuration.hpp:49-74), column by column."""
        q, v = _cols(q, self.nq), _cols(v, self.nv)
        B = q.shape[1]
        out = np.empty((self.nq, B), order="F")
        _lib().oracle_integrate(self._h, _p(q), _p(v), _p(out), ctypes.c_int64(B), int(long_double))
        return out

    ALGOS = {"rnea": 0, "aba": 1, "crba_world": 2, "crba_local": 3, "rnea_derivatives": 4, "aba_derivatives": 5}

    def count_flops(self, algo: str, q, v, a):
        """Exact operation counts of one evaluation: dict(add, mul, div, sqrt, sincos, flops)."""
        out = (ctypes.c_uint64 * 5)()
        q, v, a = (np.ascontiguousarray(x, dtype=np.float64) for x in (q, v, a))
        _lib().oracle_count_flops(self._h, self.ALGOS[algo], _p(q), _p(v), _p(a), out)
        d = dict(zip(("add", "mul", "div", "sqrt", "sincos"), (int(x) for x in out)))
        d["flops"] = d["add"] + d["mul"] + d["div"] + d["sqrt"]
        return d
