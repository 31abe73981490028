"""Per-model code generation (brbd_codegen_source): the CUDA source of a kernel specialised for one model.
Mirror of the reference's pinocchio.codegen entry points (include/pinocchio/codegen/code-generator-algo.hpp:22-570)
for the batched path."""
from __future__ import annotations

import ctypes

from . import _capi

ALGOS = {"rnea": 0, "aba": 1, "crba": 2, "rnea_derivatives": 3, "aba_derivatives": 4}


class CodegenInfo(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("record_slots", "park_slots", "nodes", "live_nodes", "adds", "muls", "recips",
                                               "sqrts", "sincos", "loads", "stores", "threads_per_block", "smem_slots", "tmem_slots", "dynamic_smem_bytes", "copies")]


def codegen_source(model, algo: str, explicit_slots: bool = False, host: bool = False, nt: int = 0, minb: int = 0,
                   direct_io: bool = False, fp32: bool = False, crba_compact: bool = False, crba_group: int = 0, crba_bulk: bool = False,
                   crba_nbuf: int = 1):
    """Returns (source text, info dict) for `model` (a Model or a flat dict) and `algo` in {"rnea", "aba"}."""
    L = _capi.lib()
    flat = model.flat() if hasattr(model, "flat") else model
    fm, keep = _capi.make_flat(flat)
    h = ctypes.c_void_p()
    _capi.check(L.brbd_model_create(ctypes.byref(fm), ctypes.byref(h)))
    try:
        flags = (1 if explicit_slots else 0) | (2 if host else 0) | (4 if fp32 else 0) | (8 if direct_io else 0) | ((nt & 0xfff) << 8) | ((minb & 0xf) << 20) \
            | (16 if crba_compact else 0) | ((crba_group & 0x1f) << 24) | ((32 | (((crba_nbuf - 1) & 3) << 29)) if crba_bulk else 0)
        src = ctypes.c_char_p()
        info = CodegenInfo()
        L.brbd_codegen_source.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(CodegenInfo)]
        L.brbd_codegen_free.argtypes = [ctypes.c_void_p]
        L.brbd_codegen_free.restype = None
        raw = ctypes.c_void_p()
        L.brbd_codegen_source.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(CodegenInfo)]
        _capi.check(L.brbd_codegen_source(h, ALGOS[algo], flags, ctypes.byref(raw), ctypes.byref(info)))
        text = ctypes.string_at(raw.value).decode()
        L.brbd_codegen_free(raw)
        return text, {k: getattr(info, k) for k, _ in CodegenInfo._fields_}
    finally:
        L.brbd_model_destroy(h)
