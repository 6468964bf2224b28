"""CPU test: the any-size path EXECUTED on the host through the library's own planner.

tests/host/generic_emul.cpp takes radices, thread geometry, shared-memory layout and the four-step split from
fft_b200/csrc/generic_plan.h (the code ssfft.cu plans with), runs generic_fft_kernel and the stand-alone RealFFT passes as
fibers (tests/host/simt/) in the same launch sequence as the library, and compares with the oracle: the reference's
test sizes (tests/00-fft.cpp:8-16), primes with and without a codelet, mixed radices, lengths beyond one CTA (generic
four-step with the two-table epilogue twiddle), RealFFT and ModifiedRealFFT forward / inverse, both precisions.

Lengths with a large prime factor p are compared with a long-double DFT instead: the reference evaluates the roots of
its O(p^2) step on a phase rounded to V (signalsmith-fft.h:204), so for p = 1031 its own float output is off by 5e-5
(relative L2) while the kernels, with exact-phase roots, are at 6e-7 -- see DESIGN.md section 5.
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tests", "host")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_generic_path_runs_on_cpu(tmp_path, oracle):
    exe = str(tmp_path / "generic_emul")
    subprocess.run(["g++", "-std=c++17", "-O1", "-D__CUDACC__", "-DSSFFT_EMUL", "-I" + os.path.join(HOST, "simt"),
                    os.path.join(HOST, "generic_emul.cpp"), "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe], check=True, capture_output=True, timeout=900)
    parts = min(8, os.cpu_count() or 1)
    with ThreadPoolExecutor(parts) as pool:
        results = list(pool.map(lambda i: subprocess.run([exe, str(i), str(parts)], capture_output=True, text=True, timeout=1200),
                                range(parts)))
    runs = 0
    for res in results:
        assert res.returncode == 0 and "GENERIC-EMUL-OK" in res.stdout, res.stdout[-4000:] + res.stderr[-2000:]
        runs += int(res.stdout.split(" runs,")[0].split()[-1])
    assert runs >= 350
