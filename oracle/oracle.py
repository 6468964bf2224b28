"""ctypes loader for the parity oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (fft_b200/) never does.

Two back ends with one calling convention:
  * ``port``      -- oracle/liboracle.so, the plain-C restatement of signalsmith-fft.h (oracle_fft.c)
  * ``reference`` -- oracle/_ref/libssfft_ref.so, the unmodified reference header behind ref_shim.cpp
                     (present when it was built where /root/reference exists; travels with the repo)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libssfft_ref.so")

KIND_C2C_FWD, KIND_C2C_INV, KIND_R2C, KIND_C2R, KIND_MR2C, KIND_MC2R = range(6)

_port = None
_ref = None


def build(quiet: bool = True) -> None:
    """Compile the checker (make -C oracle).  Building the checker is not using it."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _sigs(lib, prefix):
    fn = getattr(lib, prefix + "_batch")
    fn.restype = ctypes.c_double
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    for name in ("fft_size_minimum", "fft_size_maximum", "realfft_size_minimum", "realfft_size_maximum"):
        f = getattr(lib, f"{prefix}_{name}")
        f.restype = ctypes.c_size_t
        f.argtypes = [ctypes.c_size_t]


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build()
        lib = ctypes.CDLL(PORT_SO)
        _sigs(lib, "oracle")
        for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            f = getattr(lib, f"oracle_fill_uniform_{suf}")
            f.restype = None
            f.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64]
            for name in ("oracle_plan_create", "oracle_rplan_create"):
                g = getattr(lib, f"{name}_{suf}")
                g.restype = ctypes.c_void_p
            getattr(lib, f"oracle_plan_create_{suf}").argtypes = [ctypes.c_size_t]
            getattr(lib, f"oracle_rplan_create_{suf}").argtypes = [ctypes.c_size_t, ctypes.c_int]
            for name in ("oracle_plan_destroy", "oracle_rplan_destroy"):
                g = getattr(lib, f"{name}_{suf}")
                g.restype = None
                g.argtypes = [ctypes.c_void_p]
            for name in ("num_factors", "num_steps", "num_twiddles"):
                g = getattr(lib, f"oracle_plan_{name}_{suf}")
                g.restype = ctypes.c_size_t
                g.argtypes = [ctypes.c_void_p]
            g = getattr(lib, f"oracle_plan_factor_{suf}")
            g.restype = ctypes.c_size_t
            g.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
            g = getattr(lib, f"oracle_plan_step_{suf}")
            g.restype = None
            g.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
            g = getattr(lib, f"oracle_plan_permutation_{suf}")
            g.restype = None
            g.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            g = getattr(lib, f"oracle_rplan_size_{suf}")
            g.restype = ctypes.c_size_t
            g.argtypes = [ctypes.c_void_p]
        _port = lib
    return _port


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def reference():
    global _ref
    if _ref is None:
        if not have_reference():
            raise FileNotFoundError(f"{REF_SO} not built (needs /root/reference at build time)")
        lib = ctypes.CDLL(REF_SO)
        _sigs(lib, "ref")
        for name in ("ref_realfft_setsize_return", "ref_realfft_size"):
            f = getattr(lib, name)
            f.restype = ctypes.c_size_t
            f.argtypes = [ctypes.c_size_t]
        lib.ref_hardware_threads.restype = ctypes.c_int
        _ref = lib
    return _ref


# --------------------------------------------------------------------------------------------
# synthetic inputs: numpy twin of oracle_fill_uniform_* (and of the CUDA generator in the product)
# --------------------------------------------------------------------------------------------
def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def uniform(count: int, seed: int, dtype=np.float32, first_idx: int = 0) -> np.ndarray:
    """count scalars i.i.d. uniform in [-0.5, 0.5) from the shared counter-based generator."""
    with np.errstate(over="ignore"):
        idx = np.arange(first_idx, first_idx + count, dtype=np.uint64) + (np.uint64(seed) << np.uint64(40))
    h = _splitmix64(idx)
    if np.dtype(dtype) == np.float32:
        return ((h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0) - np.float32(0.5))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5


def uniform_complex(shape, seed: int, dtype=np.complex64, first_idx: int = 0) -> np.ndarray:
    rdt = np.float32 if np.dtype(dtype) == np.complex64 else np.float64
    n = int(np.prod(shape))
    return uniform(2 * n, seed, rdt, first_idx).view(dtype).reshape(shape)


# --------------------------------------------------------------------------------------------
# batched transforms
# --------------------------------------------------------------------------------------------
def _run(lib, prefix, kind, x, n, threads):
    x = np.ascontiguousarray(x)
    if x.dtype in (np.complex64, np.float32):
        prec = 0
    elif x.dtype in (np.complex128, np.float64):
        prec = 1
    else:
        raise TypeError(x.dtype)
    rdt = np.float32 if prec == 0 else np.float64
    cdt = np.complex64 if prec == 0 else np.complex128
    if kind in (KIND_C2C_FWD, KIND_C2C_INV):
        assert x.dtype == cdt and x.shape[-1] == n
        batch = x.size // n if n else 0
        out = np.empty_like(x)
    elif kind in (KIND_R2C, KIND_MR2C):
        nr = (n // 2) * 2
        assert x.dtype == rdt and x.shape[-1] == nr
        batch = x.size // nr if nr else 0
        out = np.zeros(x.shape[:-1] + (nr // 2,), dtype=cdt)
    else:
        nr = (n // 2) * 2
        assert x.dtype == cdt and x.shape[-1] == nr // 2
        batch = x.size // (nr // 2) if nr else 0
        out = np.zeros(x.shape[:-1] + (nr,), dtype=rdt)
    if batch == 0 or n == 0:
        return out, 0.0
    secs = getattr(lib, prefix + "_batch")(kind, prec, n, batch, x.ctypes.data, out.ctypes.data, threads)
    if secs < 0:
        raise RuntimeError("oracle batch failed")
    return out, secs


def run(kind, x, n, threads: int = 1, impl: str = "port"):
    """Transform a batch (last axis = transform).  Returns (out, seconds_of_slowest_thread)."""
    if impl == "port":
        return _run(port(), "oracle", kind, x, n, threads)
    if impl == "reference":
        return _run(reference(), "ref", kind, x, n, threads)
    raise ValueError(impl)


def fft(x, impl="port", threads=1):
    return run(KIND_C2C_FWD, x, x.shape[-1], threads, impl)[0]


def ifft(x, impl="port", threads=1):
    return run(KIND_C2C_INV, x, x.shape[-1], threads, impl)[0]


def rfft(x, modified=False, impl="port", threads=1):
    return run(KIND_MR2C if modified else KIND_R2C, x, x.shape[-1], threads, impl)[0]


def irfft(x, modified=False, impl="port", threads=1):
    return run(KIND_MC2R if modified else KIND_C2R, x, 2 * x.shape[-1], threads, impl)[0]


def rel_l2(y, ref) -> float:
    """max over the batch of ||y - ref||_2 / ||ref||_2 (SURVEY.md section 8d parity metric)."""
    y = np.asarray(y)
    ref = np.asarray(ref)
    if y.size == 0:
        return 0.0
    y2 = y.reshape(-1, y.shape[-1]).astype(np.complex128 if np.iscomplexobj(y) else np.float64)
    r2 = ref.reshape(-1, ref.shape[-1]).astype(y2.dtype)
    num = np.sqrt((np.abs(y2 - r2) ** 2).sum(axis=-1))
    den = np.sqrt((np.abs(r2) ** 2).sum(axis=-1))
    den = np.where(den == 0, 1.0, den)
    return float((num / den).max())


def plan_info(n: int, prec: str = "f64"):
    """factors, steps [(type, factor, start, inner, outer)], #twiddles, permutation (dest per source)."""
    lib = port()
    p = getattr(lib, f"oracle_plan_create_{prec}")(n)
    try:
        nf = getattr(lib, f"oracle_plan_num_factors_{prec}")(p)
        factors = [getattr(lib, f"oracle_plan_factor_{prec}")(p, i) for i in range(nf)]
        ns = getattr(lib, f"oracle_plan_num_steps_{prec}")(p)
        steps = []
        buf = (ctypes.c_size_t * 5)()
        for i in range(ns):
            getattr(lib, f"oracle_plan_step_{prec}")(p, i, buf)
            steps.append(tuple(int(v) for v in buf))
        ntw = getattr(lib, f"oracle_plan_num_twiddles_{prec}")(p)
        perm = np.zeros(max(n, 1), dtype=np.uint64)
        getattr(lib, f"oracle_plan_permutation_{prec}")(p, perm.ctypes.data)
        return factors, steps, int(ntw), perm[:n].astype(np.int64)
    finally:
        getattr(lib, f"oracle_plan_destroy_{prec}")(p)
