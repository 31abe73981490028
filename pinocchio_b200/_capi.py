"""ctypes binding of the C ABI declared in include/pinocchio_b200.h.

The shared library is built in-tree (pinocchio_b200/csrc/libpinocchio_b200.so, `make -C
pinocchio_b200/csrc` or __graft_entry__.build()).  There is no CPU fallback: if the library is
missing, or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpinocchio_b200.so")

BRBD_OK, BRBD_EINVAL, BRBD_EUNSUPPORTED_JOINT, BRBD_ETOPOLOGY, BRBD_ECUDA, BRBD_ENOMEM = range(6)
BRBD_PTR_HOST, BRBD_PTR_DEVICE, BRBD_FP64, BRBD_FP32, BRBD_ASYNC = 0, 1, 0, 2, 4

# every symbol include/pinocchio_b200.h declares
SYMBOLS = [
    "brbd_last_error_string", "brbd_version", "brbd_device_count", "brbd_model_create", "brbd_model_destroy",
    "brbd_model_nq", "brbd_model_nv", "brbd_model_njoints", "brbd_pool_create", "brbd_pool_destroy",
    "brbd_pool_size", "brbd_pool_update", "brbd_pool_set_stream", "brbd_pool_synchronize",
    "brbd_pool_launch_count", "brbd_pool_last_kernel_ms", "brbd_rnea_batch", "brbd_aba_batch", "brbd_crba_batch",
    "brbd_rnea_derivatives_batch", "brbd_aba_derivatives_batch", "brbd_measure_fp64_peak",
    "brbd_host_register", "brbd_host_unregister", "brbd_nle_batch", "brbd_gravity_batch", "brbd_minverse_batch",
    "brbd_integrate_batch", "brbd_aba_euler_step_batch", "brbd_model_get_flat", "brbd_pool_resize", "brbd_pool_model",
    "brbd_pool_device_id", "brbd_pool_workspace_bytes", "brbd_codegen_source", "brbd_codegen_free", "brbd_pool_specialize",
    "brbd_pool_specialized", "brbd_pool_set_specialized_min_batch", "brbd_model_from_urdf",
    "brbd_crba_packed_batch", "brbd_model_crba_pattern", "brbd_pool_set_host_threads",
    "brbd_crba_expand_packed",
]


class FlatModel(ctypes.Structure):
    """struct brbd_flat_model."""
    _fields_ = [
        ("njoints", ctypes.c_int32), ("nq", ctypes.c_int32), ("nv", ctypes.c_int32),
        ("parents", ctypes.POINTER(ctypes.c_int32)), ("joint_type", ctypes.POINTER(ctypes.c_int32)),
        ("idx_q", ctypes.POINTER(ctypes.c_int32)), ("idx_v", ctypes.POINTER(ctypes.c_int32)),
        ("placement", ctypes.POINTER(ctypes.c_double)), ("inertia", ctypes.POINTER(ctypes.c_double)),
        ("armature", ctypes.POINTER(ctypes.c_double)), ("gravity", ctypes.c_double * 3),
        ("axis", ctypes.POINTER(ctypes.c_double)),
    ]


class EngineError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[brbd status {status}] {message}")
        self.status = status


_lib = None


def lib():
    """Load libpinocchio_b200.so (raises if it has not been built — no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C pinocchio_b200/csrc` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.brbd_last_error_string.restype = ctypes.c_char_p
    L.brbd_version.restype = ctypes.c_char_p
    L.brbd_device_count.restype = ci
    L.brbd_model_create.argtypes = [ctypes.POINTER(FlatModel), ctypes.POINTER(vp)]
    L.brbd_model_from_urdf.argtypes = [ctypes.c_char_p, ci, ctypes.POINTER(vp)]
    L.brbd_model_destroy.argtypes = [vp]
    L.brbd_model_destroy.restype = None
    for f in ("brbd_model_nq", "brbd_model_nv", "brbd_model_njoints"):
        getattr(L, f).argtypes = [vp]
    L.brbd_pool_create.argtypes = [vp, ctypes.POINTER(ci), ci, ctypes.POINTER(vp)]
    L.brbd_pool_destroy.argtypes = [vp]
    L.brbd_pool_destroy.restype = None
    L.brbd_pool_size.argtypes = [vp]
    L.brbd_model_get_flat.argtypes = [vp, ctypes.POINTER(FlatModel)]
    L.brbd_pool_resize.argtypes = [vp, ctypes.POINTER(ci), ci]
    L.brbd_pool_model.argtypes = [vp]
    L.brbd_pool_model.restype = vp
    L.brbd_pool_device_id.argtypes = [vp, ci]
    L.brbd_pool_workspace_bytes.argtypes = [vp, ci]
    L.brbd_pool_workspace_bytes.restype = ctypes.c_uint64
    L.brbd_pool_specialize.argtypes = [vp, ci, ci]
    L.brbd_pool_specialized.argtypes = [vp]
    L.brbd_pool_set_specialized_min_batch.argtypes = [vp, i64]
    L.brbd_pool_update.argtypes = [vp, vp]
    L.brbd_pool_set_stream.argtypes = [vp, vp]
    L.brbd_pool_synchronize.argtypes = [vp]
    L.brbd_pool_launch_count.argtypes = [vp]
    L.brbd_pool_launch_count.restype = i64
    L.brbd_pool_last_kernel_ms.argtypes = [vp]
    L.brbd_pool_last_kernel_ms.restype = ctypes.c_double
    L.brbd_rnea_batch.argtypes = [vp, vp, i64, vp, i64, vp, i64, vp, i64, i64, ci]
    L.brbd_aba_batch.argtypes = [vp, vp, i64, vp, i64, vp, i64, vp, i64, i64, ci]
    L.brbd_crba_batch.argtypes = [vp, vp, i64, vp, i64, i64, ci]
    L.brbd_crba_packed_batch.argtypes = [vp, vp, i64, vp, i64, i64, ci]
    L.brbd_pool_set_host_threads.argtypes = [vp, ci]
    L.brbd_crba_expand_packed.argtypes = [vp, vp, i64, vp, i64, i64, ci, ci]
    L.brbd_model_crba_pattern.argtypes = [vp, vp, vp, i64, ctypes.POINTER(i64)]
    L.brbd_rnea_derivatives_batch.argtypes = [vp, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, i64, ci]
    L.brbd_aba_derivatives_batch.argtypes = [vp, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, i64, ci]
    L.brbd_nle_batch.argtypes = [vp, vp, i64, vp, i64, vp, i64, i64, ci]
    L.brbd_gravity_batch.argtypes = [vp, vp, i64, vp, i64, i64, ci]
    L.brbd_minverse_batch.argtypes = [vp, vp, i64, vp, i64, i64, ci]
    L.brbd_integrate_batch.argtypes = [vp, vp, i64, vp, i64, vp, i64, i64, ci]
    L.brbd_aba_euler_step_batch.argtypes = [vp, vp, i64, vp, i64, vp, i64, ctypes.c_double, vp, i64, vp, i64, i64, ci]
    L.brbd_host_register.argtypes = [vp, ctypes.c_uint64]
    L.brbd_host_unregister.argtypes = [vp]
    L.brbd_measure_fp64_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


def check(status: int):
    if status != BRBD_OK:
        raise EngineError(status, lib().brbd_last_error_string().decode())


def make_flat(flat: dict):
    """dict of numpy arrays (Model.flat()) -> (FlatModel, keep-alive list)."""
    keep = {k: np.ascontiguousarray(flat[k]) for k in
            ("parents", "joint_type", "idx_q", "idx_v", "placement", "inertia", "armature")}
    for k in ("parents", "joint_type", "idx_q", "idx_v"):
        keep[k] = keep[k].astype(np.int32)
    for k in ("placement", "inertia", "armature"):
        keep[k] = keep[k].astype(np.float64)
    keep["axis"] = np.ascontiguousarray(flat["axis"], dtype=np.float64) if "axis" in flat else np.zeros(3 * int(flat["njoints"]))
    if keep["armature"].size == 0:
        keep["armature"] = np.zeros(1)
    fm = FlatModel()
    fm.njoints, fm.nq, fm.nv = int(flat["njoints"]), int(flat["nq"]), int(flat["nv"])
    ip = lambda k: keep[k].ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    dp = lambda k: keep[k].ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    fm.parents, fm.joint_type, fm.idx_q, fm.idx_v = ip("parents"), ip("joint_type"), ip("idx_q"), ip("idx_v")
    fm.placement, fm.inertia, fm.armature, fm.axis = dp("placement"), dp("inertia"), dp("armature"), dp("axis")
    for k in range(3):
        fm.gravity[k] = float(flat["gravity"][k])
    return fm, keep
