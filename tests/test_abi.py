"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/ssfft.h declares, its
pure-integer helpers match the oracle, and the product path refuses to run without a device
(no CPU fallback).  No compute entry point is called with real work here."""
import ctypes
import os

import numpy as np
import pytest

import fft_b200
from fft_b200 import _lib as L


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        fft_b200.build()
    return fft_b200.load()


def test_exports_every_declared_symbol(lib):
    names = fft_b200.declared_symbols()
    assert len(names) >= 20 and len(set(names)) == len(names)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ssfft.h but not exported"


def test_size_helpers_match_oracle(lib, oracle):
    port = oracle.port()
    for i in range(1, 3000):
        assert lib.ssfft_size_minimum(i) == port.oracle_fft_size_minimum(i)
        assert lib.ssfft_size_maximum(i) == port.oracle_fft_size_maximum(i)
        assert lib.ssfft_real_size_minimum(i) == port.oracle_realfft_size_minimum(i)
        assert lib.ssfft_real_size_maximum(i) == port.oracle_realfft_size_maximum(i)
    assert fft_b200.FFT.sizeMinimum(1025) == 1152 and fft_b200.FFT.sizeMaximum(1025) == 1024
    assert fft_b200.RealFFT.sizeMinimum(256) == 258 and fft_b200.RealFFT.sizeMaximum(1000) == 1024


def test_error_strings_and_version(lib):
    assert lib.ssfft_error_string(L.SSFFT_OK) == b"ok"
    assert b"no CPU fallback" in lib.ssfft_error_string(L.SSFFT_ERR_NO_DEVICE)
    assert lib.ssfft_version().startswith(b"ssfft-b200")
    assert lib.ssfft_launch_count() >= 0


def test_invalid_arguments(lib):
    plan = ctypes.c_void_p()
    assert lib.ssfft_plan_create(None, L.SSFFT_C2C, L.SSFFT_F32, 16, -1) == L.SSFFT_ERR_INVALID
    assert lib.ssfft_plan_create(ctypes.byref(plan), 7, L.SSFFT_F32, 16, -1) == L.SSFFT_ERR_INVALID
    assert lib.ssfft_plan_create(ctypes.byref(plan), L.SSFFT_C2C, 9, 16, -1) == L.SSFFT_ERR_INVALID
    assert lib.ssfft_exec_c2c(None, None, None, 1, L.SSFFT_FORWARD, None) == L.SSFFT_ERR_INVALID
    assert lib.ssfft_plan_destroy(None) == L.SSFFT_OK


def test_no_cpu_fallback_without_device(lib):
    """Without a GPU the library must fail loudly rather than compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    plan = ctypes.c_void_p()
    assert lib.ssfft_plan_create(ctypes.byref(plan), L.SSFFT_C2C, L.SSFFT_F32, 4096, -1) == L.SSFFT_ERR_NO_DEVICE
    with pytest.raises(fft_b200.SsfftError):
        fft_b200.FFT(4096)
    x = np.zeros(8, np.complex64)
    with pytest.raises(fft_b200.SsfftError):
        fft_b200.FFT(8).fft(x, np.empty_like(x))


def test_product_never_imports_the_oracle():
    """Nothing under fft_b200/ or include/ may reference oracle/ (the oracle is the checker, not the product)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for sub in ("fft_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            if "build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_ssfft_io_layout_matches_the_ctypes_mirror(tmp_path):
    """struct ssfft_io as gcc lays it out (plain C, include/ssfft.h) == fft_b200._lib.SsfftIo field by field."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields = [name for name, _ in L.SsfftIo._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ssfft.h"\nint main(void) {\n'
                   '  printf("%zu", sizeof(ssfft_io));\n' +
                   "".join(f'  printf(" %zu", offsetof(ssfft_io, {f}));\n' for f in fields) + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    import subprocess
    subprocess.run(["gcc", "-std=c99", "-pedantic-errors", "-I" + os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got[0] == ctypes.sizeof(L.SsfftIo)
    assert got[1:] == [getattr(L.SsfftIo, f).offset for f in fields]


def test_extended_calls_reject_bad_handles(lib):
    io = L.SsfftIo()
    assert lib.ssfft_exec_c2c_ex(None, None, None, 1, L.SSFFT_FORWARD, ctypes.byref(io), None) == L.SSFFT_ERR_INVALID
    assert lib.ssfft_exec_r2c_ex(None, None, None, 1, ctypes.byref(io), None) == L.SSFFT_ERR_INVALID
    assert lib.ssfft_exec_c2r_ex(None, None, None, 1, None, None) == L.SSFFT_ERR_INVALID


def test_extended_request_validation_rules(tmp_path):
    """ex_validate (fft_b200/csrc/ex_request.h, CUDA-free): defaults, kinds, alignment, overlap and in-place rules."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "test_ex_request"
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", os.path.join(root, "tests", "host", "test_ex_request.cpp"), "-o", str(exe)],
                   check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0 and "EX-REQUEST-TESTS OK" in res.stdout, res.stdout + res.stderr
