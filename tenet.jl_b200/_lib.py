"""ctypes binding of libtnb200.so — exactly what Julia's `@ccall` glue does.

There is no CPU fallback anywhere in this package: if the shared library or a B200 is missing, the first
call that needs it raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtnb200.so")

TNB_OK, TNB_EINVAL, TNB_ENOMEM, TNB_ECUDA, TNB_ENCCL, TNB_EUNSUPPORTED = range(6)
TNB_C128, TNB_C64, TNB_F64, TNB_F32 = range(4)
TNB_OPT_C64_MODE, TNB_OPT_FORCE_KERNEL, TNB_OPT_GEMM_PAIR, TNB_OPT_CUDA_GRAPH = 1, 2, 3, 4
TNB_C64_SIMT, TNB_C64_TF32X3, TNB_C64_TF32X3_FAST = 0, 1, 2
KERNEL_NAMES = {0: "generic", 1: "c64_tf32x3", 2: "c128_dmma", 3: "stream", 4: "splitk", 5: "stem", 6: "stem_tc"}

DTYPE_CODE = {np.dtype(np.complex128): TNB_C128, np.dtype(np.complex64): TNB_C64,
              np.dtype(np.float64): TNB_F64, np.dtype(np.float32): TNB_F32}
CODE_DTYPE = {v: k for k, v in DTYPE_CODE.items()}


class TnbError(RuntimeError):
    """A non-zero status from the C ABI (the Julia side turns these into `error(...)`/ArgumentError)."""

    def __init__(self, code, msg):
        super().__init__(f"tnb200 error {code}: {msg}")
        self.code = code


class tnb_tensor(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("offset_elems", C.c_int64), ("dtype", C.c_int32), ("rank", C.c_int32),
                ("extent", C.POINTER(C.c_int64)), ("stride_elems", C.POINTER(C.c_int64)),
                ("mode", C.POINTER(C.c_int32)), ("conj", C.c_int32)]


class tnb_plan_info(C.Structure):
    _fields_ = [("nslices", C.c_int64), ("nsteps_per_slice", C.c_int64), ("nsteps_hoisted", C.c_int64),
                ("flops_per_slice", C.c_double), ("flops_hoisted", C.c_double),
                ("bytes_per_slice", C.c_double), ("bytes_hoisted", C.c_double),
                ("workspace_bytes", C.c_int64), ("table_bytes", C.c_int64), ("max_intermediate_elems", C.c_int64)]


class tnb_step_info(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64), ("L", C.c_int64),
                ("kernel", C.c_int32), ("hoisted", C.c_int32), ("flops", C.c_double), ("bytes", C.c_double)]


# every symbol include/tnb200.h declares (tests check the library exports each of them)
ABI_SYMBOLS = [
    "tnb_ctx_create", "tnb_ctx_destroy", "tnb_ctx_set_option", "tnb_last_error", "tnb_sync", "tnb_ctx_stream",
    "tnb_ctx_launch_count", "tnb_ctx_last_kernel", "tnb_alloc", "tnb_free", "tnb_upload", "tnb_download", "tnb_memset_zero", "tnb_buf_ptr",
    "tnb_buf_bytes", "tnb_mem_stats", "tnb_mem_trim", "tnb_binary_einsum", "tnb_binary_einsum_result",
    "tnb_plan_create", "tnb_plan_create_dry", "tnb_plan_execute", "tnb_plan_destroy", "tnb_plan_get_info",
    "tnb_plan_get_step", "tnb_plan_profile", "tnb_plan_get_step_time", "tnb_plan_dump_table", "tnb_plan_dump_step", "tnb_contract_path", "tnb_multi_contract_path", "tnb_qr_thin", "tnb_svd_thin", "tnb_comm_unique_id",
    "tnb_comm_init", "tnb_comm_allreduce_sum", "tnb_comm_size", "tnb_comm_destroy",
]

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libtnb200.so (built in-tree by `__graft_entry__.build()` / tenet.jl_b200/build.py)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "tenet.jl_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
        pt = C.POINTER(tnb_tensor)
        sig = {
            "tnb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
            "tnb_ctx_destroy": (C.c_int, [vp]),
            "tnb_ctx_set_option": (C.c_int, [vp, C.c_int, i64]),
            "tnb_last_error": (C.c_char_p, [vp]),
            "tnb_sync": (C.c_int, [vp]),
            "tnb_ctx_stream": (vp, [vp]),
            "tnb_ctx_launch_count": (i64, [vp]),
            "tnb_ctx_last_kernel": (C.c_int, [vp]),
            "tnb_alloc": (C.c_int, [vp, sz, C.POINTER(vp)]),
            "tnb_free": (C.c_int, [vp, vp]),
            "tnb_upload": (C.c_int, [vp, vp, sz, vp, sz]),
            "tnb_download": (C.c_int, [vp, vp, sz, vp, sz]),
            "tnb_memset_zero": (C.c_int, [vp, vp, sz, sz]),
            "tnb_buf_ptr": (vp, [vp]),
            "tnb_buf_bytes": (sz, [vp]),
            "tnb_mem_stats": (C.c_int, [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]),
            "tnb_mem_trim": (C.c_int, [vp]),
            "tnb_binary_einsum": (C.c_int, [vp, pt, pt, pt, C.POINTER(i32), i32, vp, vp]),
            "tnb_binary_einsum_result": (C.c_int, [pt, pt, C.POINTER(i32), i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
            "tnb_plan_create": (C.c_int, [vp, pt, i32, C.POINTER(i32), i32, C.POINTER(i32), i32, pt, C.POINTER(vp)]),
            "tnb_plan_create_dry": (C.c_int, [pt, i32, C.POINTER(i32), i32, C.POINTER(i32), i32, pt, C.POINTER(vp)]),
            "tnb_plan_execute": (C.c_int, [vp, vp, i64, i64, i64, i32]),
            "tnb_plan_destroy": (C.c_int, [vp, vp]),
            "tnb_plan_get_info": (C.c_int, [vp, C.POINTER(tnb_plan_info)]),
            "tnb_plan_get_step": (C.c_int, [vp, i32, C.POINTER(tnb_step_info)]),
            "tnb_plan_profile": (C.c_int, [vp, vp, i32]),
            "tnb_plan_get_step_time": (C.c_int, [vp, i32, C.POINTER(C.c_double), C.POINTER(i64)]),
            "tnb_plan_dump_table": (i64, [vp, i32, i32, C.POINTER(i64), i64]),
            "tnb_plan_dump_step": (C.c_int, [vp, i32, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i64), C.POINTER(i32)]),
            "tnb_contract_path": (C.c_int, [vp, pt, i32, C.POINTER(i32), i32, C.POINTER(i32), i32, i64, i64, i64, pt]),
            "tnb_multi_contract_path": (C.c_int, [vp, pt, i32, C.POINTER(i32), i32, C.POINTER(i32), i32, pt, i32]),
            "tnb_qr_thin": (C.c_int, [vp, pt, C.POINTER(i32), i32, i32, pt, pt]),
            "tnb_svd_thin": (C.c_int, [vp, pt, C.POINTER(i32), i32, i32, pt, pt, pt]),
            "tnb_comm_unique_id": (C.c_int, [vp]),
            "tnb_comm_init": (C.c_int, [vp, vp, i32, i32]),
            "tnb_comm_allreduce_sum": (C.c_int, [vp, vp, sz, i64, i32]),
            "tnb_comm_size": (i32, [vp]),
            "tnb_comm_destroy": (C.c_int, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(ctx_handle, rc):
    if rc != TNB_OK:
        msg = load_library().tnb_last_error(ctx_handle)
        raise TnbError(rc, (msg or b"").decode("utf-8", "replace"))


def make_desc(buf_handle, offset, dtype_code, extents, strides, modes, conj=False):
    """Build a tnb_tensor; returns (struct, keepalive) — keepalive owns the ctypes arrays."""
    r = len(extents)
    ext = (C.c_int64 * max(r, 1))(*[int(x) for x in extents])
    st = (C.c_int64 * max(r, 1))(*[int(x) for x in strides])
    md = (C.c_int32 * max(r, 1))(*[int(x) for x in modes])
    t = tnb_tensor(buf_handle, int(offset), int(dtype_code), r, ext, st, md, 1 if conj else 0)
    return t, (ext, st, md)
