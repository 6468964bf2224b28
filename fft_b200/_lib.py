"""ctypes binding of the C ABI declared in include/ssfft.h.

Fails loudly when libssfft.so is missing: there is no CPU fallback anywhere in the product path.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssfft.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ssfft.h")

SSFFT_OK, SSFFT_ERR_INVALID, SSFFT_ERR_CUDA, SSFFT_ERR_UNSUPPORTED, SSFFT_ERR_NO_DEVICE, SSFFT_ERR_ALLOC = range(6)
SSFFT_C2C, SSFFT_REAL, SSFFT_REAL_MODIFIED = 0, 1, 2
SSFFT_F32, SSFFT_F64 = 0, 1
SSFFT_FORWARD, SSFFT_INVERSE = 1, -1
SSFFT_MUL_NONE, SSFFT_MUL_REAL, SSFFT_MUL_COMPLEX = 0, 1, 2


class SsfftIo(ctypes.Structure):
    """struct ssfft_io (include/ssfft.h): layouts and fused multipliers of the extended execution calls."""
    _fields_ = [("in_stride", ctypes.c_int64), ("in_dist", ctypes.c_int64),
                ("out_stride", ctypes.c_int64), ("out_dist", ctypes.c_int64),
                ("pre", ctypes.c_void_p), ("pre_kind", ctypes.c_int32), ("post_kind", ctypes.c_int32),
                ("pre_dist", ctypes.c_int64), ("post", ctypes.c_void_p), ("post_dist", ctypes.c_int64)]


_lib = None


class SsfftError(RuntimeError):
    def __init__(self, status: int, where: str):
        lib = load()
        msg = lib.ssfft_error_string(status).decode()
        detail = lib.ssfft_last_cuda_error().decode()
        super().__init__(f"{where}: {msg}" + (f" [{detail}]" if detail and status == SSFFT_ERR_CUDA else ""))
        self.status = status


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into fft_b200/libssfft.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], check=True, stdout=out)
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every SSFFT_API function declared in include/ssfft.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return re.findall(r"SSFFT_API\s+[\w\s\*]+?\b(ssfft_\w+)\s*\(", text)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(fft_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, sz, i32, u64 = c.c_void_p, c.c_size_t, c.c_int, c.c_uint64
    sig = {
        "ssfft_size_minimum": (sz, [sz]), "ssfft_size_maximum": (sz, [sz]),
        "ssfft_real_size_minimum": (sz, [sz]), "ssfft_real_size_maximum": (sz, [sz]),
        "ssfft_plan_create": (i32, [c.POINTER(vp), i32, i32, sz, i32]),
        "ssfft_plan_destroy": (i32, [vp]),
        "ssfft_plan_size": (sz, [vp]),
        "ssfft_plan_describe": (i32, [vp, c.c_char_p, sz]),
        "ssfft_exec_c2c": (i32, [vp, vp, vp, sz, i32, vp]),
        "ssfft_exec_r2c": (i32, [vp, vp, vp, sz, vp]),
        "ssfft_exec_c2r": (i32, [vp, vp, vp, sz, vp]),
        "ssfft_exec_c2c_ex": (i32, [vp, vp, vp, sz, i32, c.POINTER(SsfftIo), vp]),
        "ssfft_exec_r2c_ex": (i32, [vp, vp, vp, sz, c.POINTER(SsfftIo), vp]),
        "ssfft_exec_c2r_ex": (i32, [vp, vp, vp, sz, c.POINTER(SsfftIo), vp]),
        "ssfft_exec_host": (i32, [vp, i32, vp, vp, sz]),
        "ssfft_device_count": (i32, [c.POINTER(i32)]),
        "ssfft_malloc": (i32, [c.POINTER(vp), sz]),
        "ssfft_free": (i32, [vp]),
        "ssfft_memcpy_h2d": (i32, [vp, vp, sz, vp]),
        "ssfft_memcpy_d2h": (i32, [vp, vp, sz, vp]),
        "ssfft_stream_synchronize": (i32, [vp]),
        "ssfft_fill_uniform": (i32, [vp, sz, i32, u64, u64, vp]),
        "ssfft_transpose_twiddle": (i32, [vp, vp, sz, sz, sz, sz, u64, i32, i32, vp]),
        "ssfft_permute102": (i32, [vp, vp, sz, sz, sz, i32, vp]),
        "ssfft_ipc_export": (i32, [vp, vp]),
        "ssfft_ipc_import": (i32, [vp, c.POINTER(vp)]),
        "ssfft_ipc_close": (i32, [vp]),
        "ssfft_exchange_transpose": (i32, [vp, c.POINTER(vp), i32, sz, sz, sz, sz, sz, u64, i32, i32, vp]),
        "ssfft_memcpy_d2d": (i32, [vp, vp, sz, vp]),
        "ssfft_host_alloc": (i32, [c.POINTER(vp), sz]),
        "ssfft_host_free": (i32, [vp]),
        "ssfft_dist_plan_create": (i32, [c.POINTER(vp), i32, sz, i32, c.POINTER(i32), i32]),
        "ssfft_dist_plan_destroy": (i32, [vp]),
        "ssfft_dist_plan_describe": (i32, [vp, c.c_char_p, sz]),
        "ssfft_dist_plan_factor": (sz, [vp, i32]),
        "ssfft_dist_exec_c2c": (i32, [vp, c.POINTER(vp), c.POINTER(vp), i32]),
        "ssfft_dist_synchronize": (i32, [vp]),
        "ssfft_dist_wait": (i32, [vp, i32, vp]),
        "ssfft_error_string": (c.c_char_p, [i32]),
        "ssfft_last_cuda_error": (c.c_char_p, []),
        "ssfft_launch_count": (u64, []),
        "ssfft_version": (c.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, where: str) -> None:
    if status != SSFFT_OK:
        raise SsfftError(status, where)
