"""CPU test: the ticket-queue four-step kernels (fft_b200/csrc/flat.cuh) EXECUTED on the host for every registered pair.

tests/host/flat_emul.cpp compiles the unmodified kernel source with g++ against tests/host/simt/: consumer threads with
their named barrier, the TMA producer thread with its ticket counter, the dependency counters in global memory and the
full / empty / done mbarriers of the shared-memory ring all run as fibers; a few CTAs are resident and scheduled under
three skews, both fiber orders and both ends of the TMA copies' legal timing; `discard.global.L2` poisons what it drops.
The pair list is read from the registry sources (flat_f*.cu).  Every pair runs C2C forward and inverse with ring depth
1 and 2 over four (CTAs, delay, scratch slots) shapes against the oracle.
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

_PAIR = re.compile(r'make_flat_entry<TileCfg<([^>]*)>,\s*TileCfg<([^>]*)>,\s*(\d+),\s*(\d+)(?:,\s*\w+)*>\("([^"]*)"\)')


_ALIAS = re.compile(r'using\s+(\w+)\s*=\s*(TileCfg<[^>]*>)\s*;')


def registered_pairs():
    out = {}
    for name in sorted(os.listdir(CSRC)):
        if not re.match(r"flat_f(32|64)_[a-z]\.cu$", name):
            continue
        aliases = {}
        for line in open(os.path.join(CSRC, name)):
            line = line.split("//")[0]
            a = _ALIAS.search(line)
            if a:
                aliases[a.group(1)] = a.group(2)
                continue
            if "push_back" not in line:
                continue
            for alias, full in aliases.items():  # `using A256 = TileCfg<...>;` names used in the entries
                line = re.sub(r"\b%s\b" % alias, full, line)
            m = _PAIR.search(line)
            assert m, f"unparsed registry line in {name}: {line}"
            out.setdefault((m.group(1), m.group(2)), m.group(5))  # ring depth / CTAs per SM do not matter on the CPU
    return [(a, b, n) for (a, b), n in out.items()]


def test_pair_registry_is_parsed():
    names = {p[2].rsplit("_", 1)[0] for p in registered_pairs()}
    assert "float_flat_256x256" in names and "float_flat_128x256" in names


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_every_flat_kernel_runs_on_cpu(tmp_path, oracle):
    pairs = registered_pairs()
    if not os.environ.get("SSFFT_EMUL_ALL"):
        # the pairs with a 2048-point leg take minutes each on the CPU: they share every code path with the pairs below
        # (same kernel template, three-pass tiles, in-place ring of one) and run with SSFFT_EMUL_ALL=1
        pairs = [p for p in pairs if int(p[0].split(",")[1]) * int(p[1].split(",")[1]) <= (1 << 20)]
    pairs.sort(key=lambda p: -int(p[0].split(",")[1]) * int(p[1].split(",")[1]))
    nchunks = min(8, os.cpu_count() or 1, len(pairs))
    chunks = [pairs[i::nchunks] for i in range(nchunks)]

    def build_and_run(i):
        inc = tmp_path / f"pairs_{i}.inc"
        inc.write_text("".join(f'PAIR(({a}), ({b}), "{n}")\n' for a, b, n in chunks[i]))
        exe = str(tmp_path / f"flat_emul_{i}")
        cmd = ["g++", "-std=c++17", "-O1", "-D__CUDACC__", "-DSSFFT_EMUL", f'-DFLAT_PAIR_INC="{inc}"',
               "-I" + os.path.join(HOST, "simt"), os.path.join(HOST, "flat_emul.cpp"),
               "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe]
        subprocess.run(cmd, check=True, capture_output=True, timeout=900)
        return subprocess.run([exe], capture_output=True, text=True, timeout=1200)

    with ThreadPoolExecutor(nchunks) as pool:
        results = list(pool.map(build_and_run, range(nchunks)))
    runs = 0
    for res in results:
        assert res.returncode == 0 and "FLAT-EMUL-OK" in res.stdout, res.stdout[-4000:] + res.stderr[-2000:]
        runs += int(res.stdout.split(" runs,")[0].split()[-1])
    # (ring 1, 2 separate exchange buffer + ring 1, 2, 3 in place) x 4 shapes x forward / inverse, + RealFFT forward / inverse
    # x 3 shapes x ring 1, 2 for the pairs that carry the real kernels
    assert runs >= 16 * len(pairs)  # lengths from 2^18 on run two ring variants (+ one real), wide tiles skip rings that do not fit
