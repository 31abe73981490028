"""Host-side mirror of the reference's batched interface, on top of the C ABI.

Mirrors, with the same names, argument order and error behaviour:

* ``ModelPool``               — ModelPoolTpl (include/pinocchio/multibody/pool/model.hpp:19-165;
                                Python binding include/pinocchio/bindings/python/multibody/pool/model.hpp:30-80)
* ``rneaInParallel``          — include/pinocchio/algorithm/parallel/rnea.hpp:38-83
                                (binding bindings/python/algorithm/parallel/rnea.cpp:15-64)
* ``abaInParallel``           — include/pinocchio/algorithm/parallel/aba.hpp:40-84
* ``crbaInParallel``, ``computeRNEADerivativesInParallel``, ``computeABADerivativesInParallel`` —
  batched analogues of crba (algorithm/crba.hpp:47-51), computeRNEADerivatives
  (rnea-derivatives.hpp:110-128) and computeABADerivatives (aba-derivatives.hpp:52-66).

Inputs are (rows x B) arrays whose COLUMNS are configurations: numpy arrays (host; staged over PCIe
by the engine) or torch CUDA tensors of shape (B, rows) / Fortran-like (rows, B) views (device;
zero-copy).  ``num_threads`` is accepted for signature compatibility and ignored by the GPU path.
Errors the reference raises as ``std::invalid_argument`` surface as ``ValueError``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import BRBD_ASYNC, BRBD_FP32, BRBD_FP64, BRBD_PTR_DEVICE, BRBD_PTR_HOST, EngineError

try:  # torch is optional plumbing (device memory / streams); numpy inputs work without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _raise(e: EngineError):
    if e.status in (_capi.BRBD_EINVAL, _capi.BRBD_EUNSUPPORTED_JOINT, _capi.BRBD_ETOPOLOGY):
        raise ValueError(str(e)) from None
    raise e


class ModelPool:
    """Device analogue of ``pinocchio.ModelPool``: one staged model replica per GPU."""

    def __init__(self, model, devices: Optional[Sequence[int]] = None):
        self._model = model
        self._h_model = ctypes.c_void_p()
        self._h_pool = ctypes.c_void_p()
        self._stream_pinned = False   # True once the caller chose a stream with set_stream()
        self._stream_handle = None    # what brbd_pool_set_stream was last given
        self._create_model(model)
        L = _capi.lib()
        if devices is None:
            ids, n = None, 0
        else:
            arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
            ids, n = arr, len(devices)
        try:
            _capi.check(L.brbd_pool_create(self._h_model, ids, n, ctypes.byref(self._h_pool)))
        except EngineError as e:
            _raise(e)

    def _create_model(self, model):
        L = _capi.lib()
        flat = model.flat() if hasattr(model, "flat") else model
        fm, keep = _capi.make_flat(flat)
        h = ctypes.c_void_p()
        try:
            _capi.check(L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)))
        except EngineError as e:
            _raise(e)
        if self._h_model:
            L.brbd_model_destroy(self._h_model)
        self._h_model = h
        self.nq, self.nv, self.njoints = L.brbd_model_nq(h), L.brbd_model_nv(h), L.brbd_model_njoints(h)

    # -- ModelPoolTpl interface -----------------------------------------------------------------
    def size(self) -> int:
        return int(_capi.lib().brbd_pool_size(self._h_pool))

    def getModel(self, index: int = 0):
        """ModelPoolTpl::getModel (pool/model.hpp:60-74): every replica holds the same model."""
        if not (0 <= int(index) < self.size()):
            raise ValueError(f"Index greater than the size of the model vector: {index} >= {self.size()}")
        return self._model

    def getModels(self):
        """ModelPoolTpl::getModels (pool/model.hpp:77-89)."""
        return [self._model] * self.size()

    def getData(self, index: int = 0):
        """ModelPoolTpl::getData has no counterpart with per-thread Data on the device; what replica `index` owns is a CUDA
        device and its grow-only arenas."""
        if not (0 <= int(index) < self.size()):
            raise ValueError(f"Index greater than the size of the data vector: {index} >= {self.size()}")
        L = _capi.lib()
        return {"device": int(L.brbd_pool_device_id(self._h_pool, int(index))),
                "workspace_bytes": int(L.brbd_pool_workspace_bytes(self._h_pool, int(index)))}

    def getDatas(self):
        return [self.getData(i) for i in range(self.size())]

    def devices(self):
        L = _capi.lib()
        return [int(L.brbd_pool_device_id(self._h_pool, i)) for i in range(self.size())]

    def resize(self, devices: Sequence[int]):
        """ModelPoolTpl::resize (pool/model.hpp:110-131): the replicas of a device pool are devices."""
        arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        try:
            _capi.check(_capi.lib().brbd_pool_resize(self._h_pool, arr, len(devices)))
        except EngineError as e:
            _raise(e)
        self._stream_pinned, self._stream_handle = False, None

    def specialize(self, algos=("rnea", "aba"), fp32: bool = False, min_batch: Optional[int] = None):
        """Generate, compile (NVRTC) and load kernels specialised for this pool's model (brbd_pool_specialize): the tree
        unrolled, joint types resolved, the model's constants folded in.  Large batches of the listed algorithms then run
        them; results agree with the generic kernels to rounding."""
        from .codegen import ALGOS
        mask = 0
        for a in algos:
            if a in ("rnea_derivatives", "aba_derivatives") and self.nv > 16:
                continue  # these programs keep all 3 nv^2 results alive: small models only (the cooperative kernels serve the rest)
            mask |= 1 << ALGOS[a]
        try:
            _capi.check(_capi.lib().brbd_pool_specialize(self._h_pool, mask, 4 if fp32 else 0))
            if min_batch is not None:
                _capi.check(_capi.lib().brbd_pool_set_specialized_min_batch(self._h_pool, int(min_batch)))
        except EngineError as e:
            _raise(e)

    def crbaPattern(self):
        """(rows, cols) of the structural pattern of crba's result (brbd_model_crba_pattern): entry k of a packed column
        (crbaPackedInParallel) is M[rows[k], cols[k]]; column-major; everything outside it is a structural zero."""
        L = _capi.lib()
        n = ctypes.c_int64()
        _capi.check(L.brbd_model_crba_pattern(self._h_model, None, None, 0, ctypes.byref(n)))
        rows, cols = np.zeros(n.value, dtype=np.int32), np.zeros(n.value, dtype=np.int32)
        _capi.check(L.brbd_model_crba_pattern(self._h_model, rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p),
                                              n.value, ctypes.byref(n)))
        return rows, cols

    def specialized(self):
        from .codegen import ALGOS
        mask = int(_capi.lib().brbd_pool_specialized(self._h_pool))
        return [a for a, k in ALGOS.items() if mask & (1 << k)]

    def flat_model(self):
        """The model as the engine holds it (brbd_model_get_flat), as a dict of numpy arrays."""
        fm = _capi.FlatModel()
        _capi.check(_capi.lib().brbd_model_get_flat(self._h_model, ctypes.byref(fm)))
        n, nv = fm.njoints, fm.nv
        a = lambda ptr, k, dt: np.ctypeslib.as_array(ptr, shape=(k,)).astype(dt).copy() if k else np.zeros(0, dtype=dt)
        return {"njoints": n, "nq": fm.nq, "nv": nv, "parents": a(fm.parents, n, np.int32), "joint_type": a(fm.joint_type, n, np.int32),
                "idx_q": a(fm.idx_q, n, np.int32), "idx_v": a(fm.idx_v, n, np.int32),
                "placement": a(fm.placement, 12 * n, np.float64).reshape(n, 12), "inertia": a(fm.inertia, 10 * n, np.float64).reshape(n, 10),
                "armature": a(fm.armature, nv, np.float64), "axis": a(fm.axis, 3 * n, np.float64),
                "gravity": np.array([fm.gravity[0], fm.gravity[1], fm.gravity[2]])}

    def update(self, model):
        """ModelPoolTpl::update (pool/model.hpp:100-108)."""
        self._model = model
        self._create_model(model)
        _capi.check(_capi.lib().brbd_pool_update(self._h_pool, self._h_model))

    # -- engine extras --------------------------------------------------------------------------
    def synchronize(self):
        _capi.check(_capi.lib().brbd_pool_synchronize(self._h_pool))

    def set_stream(self, cuda_stream: Optional[int]):
        """Pin device-pointer calls to one CUDA stream (a cudaStream_t handle; torch: `stream.cuda_stream`).  A handle of 0 is
        the legacy default stream (what torch's default stream is).  `None` un-pins: torch tensors then run on torch's
        current stream of their device, raw device pointers on the pool's own stream."""
        if cuda_stream is None:
            self._stream_pinned = False
            self._set_stream_handle(None)
            return
        self._stream_pinned = True
        self._set_stream_handle(int(cuda_stream))

    _CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy

    def _set_stream_handle(self, handle: Optional[int]):
        if handle == self._stream_handle and handle is not None:
            return
        raw = 0 if handle is None else (self._CUDA_STREAM_LEGACY if handle == 0 else handle)
        _capi.check(_capi.lib().brbd_pool_set_stream(self._h_pool, ctypes.c_void_p(raw)))
        self._stream_handle = handle

    def launch_count(self) -> int:
        return int(_capi.lib().brbd_pool_launch_count(self._h_pool))

    def last_kernel_ms(self) -> float:
        return float(_capi.lib().brbd_pool_last_kernel_ms(self._h_pool))

    def measure_fp64_peak(self):
        f, ms = ctypes.c_double(), ctypes.c_double()
        _capi.check(_capi.lib().brbd_measure_fp64_peak(self._h_pool, ctypes.byref(f), ctypes.byref(ms)))
        return f.value, ms.value

    def close(self):
        L = _capi._lib
        if L is None:
            return
        if self._h_pool:
            L.brbd_pool_destroy(self._h_pool)
            self._h_pool = ctypes.c_void_p()
        if self._h_model:
            L.brbd_model_destroy(self._h_model)
            self._h_model = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# argument marshalling
# ------------------------------------------------------------------------------------------------
class _Arg:
    __slots__ = ("ptr", "ld", "rows", "cols", "device", "dtype", "keep", "writeback")


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _describe(x, rows: int, name: str, out: bool = False) -> _Arg:
    """Resolve pointer / leading dimension of a (rows x B) column-per-configuration block."""
    a = _Arg()
    a.rows, a.writeback = rows, None
    if _is_torch(x):
        if x.dim() == 1:
            x = x.view(1, -1) if rows != x.numel() else x.view(1, rows)
        # accepted: (B, rows) row-major [natural torch batch layout] or (rows, B) column-major view
        if x.shape[-1] == rows and x.stride(-1) == 1 and x.dim() == 2:
            a.cols, a.ld = x.shape[0], x.stride(0) if x.shape[0] > 1 else rows
        elif x.shape[0] == rows and x.stride(0) == 1 and x.dim() == 2:
            a.cols, a.ld = x.shape[1], x.stride(1) if x.shape[1] > 1 else rows
        else:
            raise ValueError(f"{name}: expected a (B, {rows}) contiguous tensor or a ({rows}, B) column-major view, "
                             f"got shape {tuple(x.shape)} strides {x.stride()}")
        a.device = x.is_cuda
        a.dtype = np.float64 if x.dtype == torch.float64 else (np.float32 if x.dtype == torch.float32 else None)
        a.ptr, a.keep = x.data_ptr(), x
        if a.dtype is None:
            raise ValueError(f"{name}: dtype must be float64 or float32")
        return a
    arr = x
    if not isinstance(arr, np.ndarray):
        arr = np.asarray(arr)
    if arr.ndim == 1:
        arr = arr.reshape(rows, 1, order="F") if arr.size == rows else arr.reshape(-1, 1)
    if arr.ndim != 2 or arr.shape[0] != rows:
        # same wording family as PINOCCHIO_CHECK_ARGUMENT_SIZE (macros.hpp:185-223)
        raise ValueError(f"wrong argument size: expected {rows} rows for {name}, got {arr.shape[0] if arr.ndim else 0}")
    if arr.dtype not in (np.float64, np.float32):
        if out:
            raise ValueError(f"{name}: output dtype must be float64 or float32")
        arr = arr.astype(np.float64)
    if arr.size > 0 and (arr.strides[0] != arr.itemsize or (arr.shape[1] > 1 and arr.strides[1] < rows * arr.itemsize)):
        if out:
            raise ValueError(f"{name}: output must be column-major (Fortran order)")
        arr = np.asfortranarray(arr)
    a.cols = arr.shape[1]
    a.ld = arr.strides[1] // arr.itemsize if arr.shape[1] > 1 else rows
    a.device, a.dtype, a.ptr, a.keep = False, arr.dtype.type, arr.ctypes.data, arr
    return a


def _alloc_like(ref, rows: int, cols: int):
    if _is_torch(ref):
        return torch.empty((cols, rows), dtype=ref.dtype, device=ref.device)
    return np.empty((rows, cols), dtype=ref.dtype if ref.dtype in (np.float64, np.float32) else np.float64, order="F")


def _call(pool: ModelPool, fn_name: str, ins, outs, async_: bool = False, mid=()):
    """ins / outs: lists of (_Arg or None); `mid`: extra scalar C arguments between the inputs and the outputs."""
    args = [a for a in ins + outs if a is not None]
    B = ins[0].cols
    for a in args:
        if a.cols != B:
            raise ValueError(f"wrong argument size: all blocks must have the same number of columns ({a.cols} != {B})")
    dev = {a.device for a in args}
    dt = {a.dtype for a in args}
    if len(dev) != 1:
        raise ValueError("mixing host and device arguments in one call is not supported")
    if len(dt) != 1:
        raise ValueError("all arguments must share one dtype (float64 or float32)")
    on_device = dev.pop()
    flags = (BRBD_PTR_DEVICE if on_device else BRBD_PTR_HOST) | (BRBD_FP32 if dt.pop() is np.float32 else BRBD_FP64)
    if async_:
        flags |= BRBD_ASYNC
    if on_device:
        # Device tensors: every block must live on the pool's device, and the kernels must be ordered after whatever
        # produced the inputs — so, unless the caller pinned a stream with set_stream(), the call runs on torch's CURRENT
        # stream of that device (as any torch op would), not on the pool's private non-blocking stream.
        devs = {a.keep.device.index for a in args if _is_torch(a.keep)}
        pool_devs = pool.devices()
        if len(pool_devs) != 1:
            raise ValueError("device tensors require a single-device pool")
        if devs and devs != {pool_devs[0]}:
            raise ValueError(f"tensors on cuda device(s) {sorted(devs)} passed to a pool on device {pool_devs[0]}")
        if not pool._stream_pinned and devs:
            pool._set_stream_handle(int(torch.cuda.current_stream(pool_devs[0]).cuda_stream))
    cargs = [pool._h_pool]
    for group in (ins, None, outs):
        if group is None:
            cargs += list(mid)
            continue
        for a in group:
            if a is None:
                cargs += [ctypes.c_void_p(0), ctypes.c_int64(0)]
            else:
                cargs += [ctypes.c_void_p(a.ptr), ctypes.c_int64(a.ld)]
    cargs += [ctypes.c_int64(B), ctypes.c_int(flags)]
    try:
        _capi.check(getattr(_capi.lib(), fn_name)(*cargs))
    except EngineError as e:
        _raise(e)


def pin_host(array: np.ndarray) -> np.ndarray:
    """Page-lock a caller-owned numpy block in place (brbd_host_register); returns it.  Host-pointer calls on
    pinned blocks run at the full PCIe rate and overlap upload / compute / download."""
    _capi.check(_capi.lib().brbd_host_register(ctypes.c_void_p(array.ctypes.data), ctypes.c_uint64(array.nbytes)))
    return array


def unpin_host(array: np.ndarray) -> None:
    _capi.check(_capi.lib().brbd_host_unregister(ctypes.c_void_p(array.ctypes.data)))


def _check_pool(num_threads: int, pool: ModelPool):
    # parallel/rnea.hpp:52-54 — the GPU pool has no per-thread replicas, so only emptiness is checked
    if pool.size() <= 0:
        raise ValueError("The pool should have at least one element")
    if int(num_threads) < 0:
        raise ValueError("num_threads must be non-negative")


def rneaInParallel(num_threads: int, pool: ModelPool, q, v, a, tau=None, async_: bool = False):
    """tau[:, i] = rnea(model, q[:, i], v[:, i], a[:, i]) — parallel/rnea.hpp:38-83."""
    _check_pool(num_threads, pool)
    aq, av, aa = _describe(q, pool.nq, "q"), _describe(v, pool.nv, "v"), _describe(a, pool.nv, "a")
    if tau is None:
        tau = _alloc_like(q, pool.nv, aq.cols)
    _call(pool, "brbd_rnea_batch", [aq, av, aa], [_describe(tau, pool.nv, "tau", out=True)], async_)
    return tau


def abaInParallel(num_threads: int, pool: ModelPool, q, v, tau, a=None, async_: bool = False):
    """a[:, i] = aba(model, q[:, i], v[:, i], tau[:, i], Convention::WORLD) — parallel/aba.hpp:40-84."""
    _check_pool(num_threads, pool)
    aq, av, at = _describe(q, pool.nq, "q"), _describe(v, pool.nv, "v"), _describe(tau, pool.nv, "tau")
    if a is None:
        a = _alloc_like(q, pool.nv, aq.cols)
    _call(pool, "brbd_aba_batch", [aq, av, at], [_describe(a, pool.nv, "a", out=True)], async_)
    return a


def crbaInParallel(num_threads: int, pool: ModelPool, q, M=None, async_: bool = False):
    """M[:, i] = vec(crba(model, q[:, i])): (nv*nv x B), upper triangle + zeros (crba.hpp:15-22)."""
    _check_pool(num_threads, pool)
    # the one place where the reference's num_threads does something here: with host blocks, only the entries inside the tree
    # sparsity cross PCIe and num_threads host threads rebuild the dense matrices (brbd_pool_set_host_threads)
    _capi.check(_capi.lib().brbd_pool_set_host_threads(pool._h_pool, int(num_threads)))
    aq = _describe(q, pool.nq, "q")
    nn = pool.nv * pool.nv
    if M is None:
        M = _alloc_like(q, nn, aq.cols)
    _call(pool, "brbd_crba_batch", [aq], [_describe(M, nn, "M", out=True)], async_)
    return M


def crbaPackedInParallel(num_threads: int, pool: ModelPool, q, P=None, async_: bool = False):
    """P[:, i] = the entries of crba(model, q[:, i]) inside the structural pattern (pool.crbaPattern()), column-major:
    (nnz x B).  Opt-in output format (brbd_crba_packed_batch): a third of the bytes of the dense block for a humanoid."""
    _check_pool(num_threads, pool)
    aq = _describe(q, pool.nq, "q")
    nnz = len(pool.crbaPattern()[0])
    if P is None:
        P = _alloc_like(q, nnz, aq.cols)
    _call(pool, "brbd_crba_packed_batch", [aq], [_describe(P, nnz, "P", out=True)], async_)
    return P


def expandPackedCrba(pool: ModelPool, P, M=None, num_threads: int = 1):
    """Dense (nv*nv x B) block from a packed host block (numpy, F-ordered) — what a caller that wants data.M back does
    (brbd_crba_expand_packed: num_threads host threads, a copy)."""
    aP = _describe(P, len(pool.crbaPattern()[0]), "P")
    if aP.device:
        raise ValueError("expandPackedCrba works on host blocks")
    nn = pool.nv * pool.nv
    if M is None:
        M = np.empty((nn, aP.cols), dtype=np.float32 if aP.dtype is np.float32 else np.float64, order="F")
    aM = _describe(M, nn, "M", out=True)
    if aM.cols != aP.cols:
        raise ValueError(f"wrong argument size: all blocks must have the same number of columns ({aM.cols} != {aP.cols})")
    try:
        _capi.check(_capi.lib().brbd_crba_expand_packed(pool._h_model, ctypes.c_void_p(aP.ptr), aP.ld, ctypes.c_void_p(aM.ptr), aM.ld, aP.cols,
                                                        int(num_threads), BRBD_FP32 if aP.dtype is np.float32 else BRBD_FP64))
    except EngineError as e:
        _raise(e)
    return M


def computeRNEADerivativesInParallel(num_threads: int, pool: ModelPool, q, v, a, dtau_dq=None, dtau_dv=None,
                                     dtau_da=None, tau=None, async_: bool = False):
    """Per column: computeRNEADerivatives (rnea-derivatives.hpp:110-128). Returns (dtau_dq, dtau_dv, dtau_da, tau)."""
    _check_pool(num_threads, pool)
    aq, av, aa = _describe(q, pool.nq, "q"), _describe(v, pool.nv, "v"), _describe(a, pool.nv, "a")
    nn, B = pool.nv * pool.nv, aq.cols
    dtau_dq = _alloc_like(q, nn, B) if dtau_dq is None else dtau_dq
    dtau_dv = _alloc_like(q, nn, B) if dtau_dv is None else dtau_dv
    dtau_da = _alloc_like(q, nn, B) if dtau_da is None else dtau_da
    tau = _alloc_like(q, pool.nv, B) if tau is None else tau
    _call(pool, "brbd_rnea_derivatives_batch", [aq, av, aa],
          [_describe(dtau_dq, nn, "dtau_dq", True), _describe(dtau_dv, nn, "dtau_dv", True),
           _describe(dtau_da, nn, "dtau_da", True), _describe(tau, pool.nv, "tau", True)], async_)
    return dtau_dq, dtau_dv, dtau_da, tau


def computeABADerivativesInParallel(num_threads: int, pool: ModelPool, q, v, tau, ddq_dq=None, ddq_dv=None,
                                    ddq_dtau=None, ddq=None, async_: bool = False):
    """Per column: computeABADerivatives (aba-derivatives.hpp:52-66). Returns (ddq_dq, ddq_dv, ddq_dtau, ddq)."""
    _check_pool(num_threads, pool)
    aq, av, at = _describe(q, pool.nq, "q"), _describe(v, pool.nv, "v"), _describe(tau, pool.nv, "tau")
    nn, B = pool.nv * pool.nv, aq.cols
    ddq_dq = _alloc_like(q, nn, B) if ddq_dq is None else ddq_dq
    ddq_dv = _alloc_like(q, nn, B) if ddq_dv is None else ddq_dv
    ddq_dtau = _alloc_like(q, nn, B) if ddq_dtau is None else ddq_dtau
    ddq = _alloc_like(q, pool.nv, B) if ddq is None else ddq
    _call(pool, "brbd_aba_derivatives_batch", [aq, av, at],
          [_describe(ddq_dq, nn, "ddq_dq", True), _describe(ddq_dv, nn, "ddq_dv", True),
           _describe(ddq_dtau, nn, "ddq_dtau", True), _describe(ddq, pool.nv, "ddq", True)], async_)
    return ddq_dq, ddq_dv, ddq_dtau, ddq


# ---- the callers' other needs on the same sweeps; no batched version exists upstream, the names follow
# ---- the reference's single-configuration functions ----------------------------------------------------------------
def nonLinearEffectsInParallel(num_threads: int, pool: ModelPool, q, v, nle=None, async_: bool = False):
    """nle[:, i] = nonLinearEffects(model, q[:, i], v[:, i]) — algorithm/rnea.hpp:105 (= rnea(q, v, 0))."""
    _check_pool(num_threads, pool)
    aq, av = _describe(q, pool.nq, "q"), _describe(v, pool.nv, "v")
    if nle is None:
        nle = _alloc_like(q, pool.nv, aq.cols)
    _call(pool, "brbd_nle_batch", [aq, av], [_describe(nle, pool.nv, "nle", out=True)], async_)
    return nle


def computeGeneralizedGravityInParallel(num_threads: int, pool: ModelPool, q, g=None, async_: bool = False):
    """g[:, i] = computeGeneralizedGravity(model, q[:, i]) — algorithm/rnea.hpp:133 (= rnea(q, 0, 0))."""
    _check_pool(num_threads, pool)
    aq = _describe(q, pool.nq, "q")
    if g is None:
        g = _alloc_like(q, pool.nv, aq.cols)
    _call(pool, "brbd_gravity_batch", [aq], [_describe(g, pool.nv, "g", out=True)], async_)
    return g


def computeMinverseInParallel(num_threads: int, pool: ModelPool, q, Minv=None, async_: bool = False):
    """Minv[:, i] = vec(computeMinverse(model, q[:, i])) — algorithm/aba.hpp:106: upper triangle of M^-1 + zeros."""
    _check_pool(num_threads, pool)
    aq = _describe(q, pool.nq, "q")
    nn = pool.nv * pool.nv
    if Minv is None:
        Minv = _alloc_like(q, nn, aq.cols)
    _call(pool, "brbd_minverse_batch", [aq], [_describe(Minv, nn, "Minv", out=True)], async_)
    return Minv


def integrateInParallel(num_threads: int, pool: ModelPool, q, v, qout=None, async_: bool = False):
    """qout[:, i] = integrate(model, q[:, i], v[:, i]) — algorithm/joint-configuration.hpp:49-74."""
    _check_pool(num_threads, pool)
    aq, av = _describe(q, pool.nq, "q"), _describe(v, pool.nv, "v")
    if qout is None:
        qout = _alloc_like(q, pool.nq, aq.cols)
    _call(pool, "brbd_integrate_batch", [aq, av], [_describe(qout, pool.nq, "qout", out=True)], async_)
    return qout


def abaEulerStepInParallel(num_threads: int, pool: ModelPool, q, v, tau, dt: float, q_next=None, v_next=None,
                           async_: bool = False):
    """One semi-implicit Euler step per column, on the device: a = aba(q, v, tau); v_next = v + dt a;
    q_next = integrate(q, dt v_next) (examples/simulation-pendulum.py:153-157).  Returns (q_next, v_next)."""
    _check_pool(num_threads, pool)
    aq, av, at = _describe(q, pool.nq, "q"), _describe(v, pool.nv, "v"), _describe(tau, pool.nv, "tau")
    q_next = _alloc_like(q, pool.nq, aq.cols) if q_next is None else q_next
    v_next = _alloc_like(q, pool.nv, aq.cols) if v_next is None else v_next
    _call(pool, "brbd_aba_euler_step_batch", [aq, av, at],
          [_describe(q_next, pool.nq, "q_next", True), _describe(v_next, pool.nv, "v_next", True)], async_,
          mid=(ctypes.c_double(float(dt)),))
    return q_next, v_next
