"""GPU tests of the header-only C++ front end (include/signalsmith-fft.h).

1. tests/host/test_header.cpp: host containers / iterators / device-pointer overloads vs the oracle.
2. The reference's OWN test suite (tests/00-fft.cpp, tests/01-real.cpp and its harness), compiled
   unmodified against our header by tests/host/build_reference_tests.sh where /root/reference exists.
   The binary travels to the GPU box; there it must report every reference test as passing.
"""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "host", "_build")


def _build_header_test():
    exe = os.path.join(BUILD, "test_header")
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "host", "test_header.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(
            ["g++", "-std=c++11", "-O1", src, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
             "-L" + os.path.join(ROOT, "fft_b200"), "-lssfft", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
             "-Wl,-rpath,$ORIGIN/../../../fft_b200", "-Wl,-rpath,$ORIGIN/../../../oracle", "-o", exe], check=True)
    return exe


def test_header_front_end(cuda_device, oracle):
    exe = _build_header_test()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "HEADER-TESTS OK" in res.stdout, res.stdout + res.stderr


def test_reference_suite_against_our_header(cuda_device):
    exe = os.path.join(BUILD, "reference_tests")
    if not os.path.exists(exe):
        pytest.skip("reference test binary not built (needs /root/reference at build time)")
    res = subprocess.run([exe, "--seed=1"], capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-3000:]
    assert "FAIL" not in out.upper().replace("FAILED: 0", ""), out[-3000:]
