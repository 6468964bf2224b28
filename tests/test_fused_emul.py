"""CPU test: every registered fused single-pass kernel (fft_b200/csrc/fused.cuh) EXECUTED on the host.

tests/host/fused_emul.cpp compiles the unmodified kernel source with g++ against tests/host/simt/ (a stand-in for the
CUDA headers) and runs each CUDA thread as a fiber: __syncthreads() is a real barrier, shared memory one host array, the
TMA bulk prefetch + mbarrier emulated at both ends of their allowed timing (copy lands at issue / at the first wait).
The configuration list is read from the registry sources (fused_f*.cu), so a newly registered size is covered without
touching this test.  Each configuration runs C2C forward / inverse / in place / from pointers that are only 8-byte
aligned, R2C, C2R and both ModifiedRealFFT flavours on a batch that makes CTAs loop and ends in a ragged group, and is
compared with the oracle within the tolerance of the GPU parity tests (1e-6 log2 N float, 1e-14 log2 N double).
The extended-I/O instantiation (ssfft_exec_*_ex: strided / overlapping layouts, fused windows and filters) runs fifteen
layouts per configuration (staged through shared memory and "lite" sides), checked for values AND for stray writes outside the requested layout; a misaligned vector
access (which x86 would tolerate and the GPU would not) aborts.

This is a check of index logic, barriers and staging hazards -- the GPU tests (-m gpu) remain the parity tests proper.
"""
import os
import re
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fft_b200", "csrc")
HOST = os.path.join(ROOT, "tests", "host")

_MACRO = re.compile(r"SSFFT_FUSED(_X|_PF|_REAL|)\(([^)]*)\)")


def registered_configs(pattern=r"fused_f(32|64)_[a-z]\.cu$"):
    """(T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, PF) for every entry of the fused registry, in source order."""
    out = []
    for name in sorted(os.listdir(CSRC)):
        if not re.match(pattern, name):
            continue
        for line in open(os.path.join(CSRC, name)):
            line = line.split("//")[0]
            if "push_back" not in line:
                continue
            m = _MACRO.search(line)
            assert m, f"unparsed registry line in {name}: {line}"
            kind, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
            if kind == "":        # SSFFT_FUSED(T, N, R0..R3, TX, FPB, MINB): default padding, no prefetch
                args += ["4", "0"]
            elif kind == "_PF":   # ... with the TMA prefetch
                args += ["4", "1"]
            assert len(args) == 11, (name, line)
            out.append(tuple(args))
    return out


def test_registry_is_parsed():
    cfgs = registered_configs()
    assert len(cfgs) >= 50
    sizes = {(c[0], int(c[1])) for c in cfgs}
    for n in (256, 1024, 4096, 8192, 1000, 2187, 3125, 6000):  # BASELINE configs C2 / C4 and the STFT sizes
        assert ("float", n) in sizes
    assert ("double", 1024) in sizes                            # BASELINE config C1


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
# the opt-in experiment builds (off by default in the library) run with SSFFT_EMUL_ALL=1: they double the minutes of this test
@pytest.mark.parametrize("variant", ["default", "l2_prefetch_hints"] if os.environ.get("SSFFT_EMUL_ALL") else ["default"])
def test_every_fused_kernel_runs_on_cpu(tmp_path, oracle, variant):
    """variant l2_prefetch_hints: the opt-in build -DSSFFT_FUSED_L2PF=1 (a later group of transforms hinted into L2),
    kept green so that it can be measured on the GPU as it is; the emulated hint reads the address it names."""
    extra = ["-DSSFFT_FUSED_L2PF=1"] if variant == "l2_prefetch_hints" else []
    cfgs = registered_configs()
    nchunks = min(8, os.cpu_count() or 1)
    chunks = [cfgs[i::nchunks] for i in range(nchunks)]

    def build_and_run(i):
        inc = tmp_path / f"cfgs_{i}.inc"
        inc.write_text("".join("CFG(" + ", ".join(c) + ")\n" for c in chunks[i]))
        exe = str(tmp_path / f"fused_emul_{i}")
        cmd = ["g++", "-std=c++17", "-O1", "-D__CUDACC__", "-DSSFFT_EMUL", *extra, f'-DFUSED_CFG_INC="{inc}"',
               *(["-DEMUL_EX_COPY"] if i == 0 else []),
               "-I" + os.path.join(HOST, "simt"), os.path.join(HOST, "fused_emul.cpp"),
               "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe]
        subprocess.run(cmd, check=True, capture_output=True, timeout=900)
        return subprocess.run([exe], capture_output=True, text=True, timeout=900)

    with ThreadPoolExecutor(nchunks) as pool:
        results = list(pool.map(build_and_run, range(nchunks)))
    runs = 0
    for res in results:
        assert res.returncode == 0 and "FUSED-EMUL-OK" in res.stdout, res.stdout[-4000:] + res.stderr[-2000:]
        runs += int(res.stdout.split(" runs over")[0].split()[-1])
    # per configuration: (8 plain runs + 2 launch shapes) x 2 (both fiber orders; prefetching kernels: both TMA timings) + 15 extended-I/O
    # layouts; + ex_copy_kernel and the column transposes
    assert runs >= 35 * len(cfgs) + 12 + 18
