"""CPU test: the four-step kernels (fft_b200/csrc/tiled.cuh) EXECUTED on the host for every registered pair of tiles.

tests/host/tiled_emul.cpp compiles the unmodified kernel source with g++ against tests/host/simt/ and keeps the whole
persistent grid alive as fibers: the CTAs of a thread-block cluster meet at an emulated hardware cluster barrier between
the column stage and the row stage, groups of clusters at the software barrier on a global counter, every CTA has its
own shared memory, and `discard.global.L2` poisons the scratch lines it drops (a line dropped before its last reader or
read again before it is rewritten corrupts the result).  The pair list is read from the registry sources
(fourstep_f*.cu).  Each pair runs C2C forward / inverse, R2C and C2R of N = N1*N2 (2^14 ... 2^20, both precisions)
through clusters of 4 and of 2 CTAs, groups of clusters, and the two-launch fallback, against the oracle.
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

_PAIR = re.compile(r'make_fourstep_entry<TileCfg<([^>]*)>,\s*TileCfg<([^>]*)>>\("([^"]*)"\)')


def registered_pairs():
    out = []
    for name in sorted(os.listdir(CSRC)):
        if not re.match(r"fourstep_f(32|64)_[a-z]\.cu$", name):
            continue
        for line in open(os.path.join(CSRC, name)):
            line = line.split("//")[0]
            if "push_back" not in line:
                continue
            m = _PAIR.search(line)
            assert m, f"unparsed registry line in {name}: {line}"
            out.append((m.group(1), m.group(2), m.group(3)))
    return out


def test_pair_registry_is_parsed():
    pairs = registered_pairs()
    names = {p[2] for p in pairs}
    assert len(pairs) >= 12
    assert "float_cluster_256x256" in names      # BASELINE config C3: RealFFT 65536 and complex 65536
    assert "float_cluster_1024x1024" in names    # 2^20


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
# the opt-in experiment builds (off by default in the library) run with SSFFT_EMUL_ALL=1: they double the minutes of this test
@pytest.mark.parametrize("variant", ["default", "l2_prefetch_hints"] if os.environ.get("SSFFT_EMUL_ALL") else ["default"])
def test_every_four_step_kernel_runs_on_cpu(tmp_path, oracle, variant):
    """variant l2_prefetch_hints: the opt-in build -DSSFFT_FOURSTEP_L2PF=1 (next column tile hinted into L2), kept
    green here so that it can be measured on the GPU as it is; the emulated prefetch READS the hinted address, so a
    hint outside the input buffer would fault under a sanitizer."""
    extra = ["-DSSFFT_FOURSTEP_L2PF=1"] if variant == "l2_prefetch_hints" else []
    pairs = registered_pairs()
    # largest pairs first so the parallel chunks finish together
    pairs.sort(key=lambda p: -int(p[0].split(",")[1]) * int(p[1].split(",")[1]))
    nchunks = min(8, os.cpu_count() or 1, len(pairs))
    chunks = [pairs[i::nchunks] for i in range(nchunks)]

    def build_and_run(i):
        inc = tmp_path / f"pairs_{i}.inc"
        inc.write_text("".join(f'PAIR(({a}), ({b}), "{n}")\n' for a, b, n in chunks[i]))
        exe = str(tmp_path / f"tiled_emul_{i}")
        cmd = ["g++", "-std=c++17", "-O1", "-D__CUDACC__", "-DSSFFT_EMUL", *extra, f'-DTILED_PAIR_INC="{inc}"',
               "-I" + os.path.join(HOST, "simt"), os.path.join(HOST, "tiled_emul.cpp"),
               "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe]
        subprocess.run(cmd, check=True, capture_output=True, timeout=900)
        return subprocess.run([exe], capture_output=True, text=True, timeout=1200)

    with ThreadPoolExecutor(nchunks) as pool:
        results = list(pool.map(build_and_run, range(nchunks)))
    runs = 0
    for res in results:
        assert res.returncode == 0 and "TILED-EMUL-OK" in res.stdout, res.stdout[-4000:] + res.stderr[-2000:]
        runs += int(res.stdout.split(" runs,")[0].split()[-1])
    assert runs == 16 * len(pairs)  # 4 transforms (C2C fwd / inv, R2C, C2R) x 4 execution shapes
