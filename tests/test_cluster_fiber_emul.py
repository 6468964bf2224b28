"""CPU test: the cluster-resident four-step kernel (fft_b200/csrc/cluster.cuh) EXECUTED on the host as fibers.

Complements tests/test_cluster_emul.py (which plays the kernel's host-device phases in a fixed order): here the kernel
itself runs -- all CTAs of a cluster alive at once, the all-to-all as remote stores into the peers' shared memory that
complete transaction bytes of the peers' mbarrier (st.async), the split-phase cluster barrier, the TMA tile prefetch --
with data landing as early and as late as the hardware allows, under both fiber orders, and with CTAs advancing together,
the lowest CTA racing ahead, and the highest CTA racing ahead (deleting the cluster wait in front of the all-to-all fails
32 of 48 runs).  Every registered configuration and every kind it can serve, against the oracle.
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

_ENTRY = re.compile(r'make_cluster_entry<ClusterCfg<([^>]*)>>\("([^"]*)",\s*(\d+)u,\s*(\d+)u\)')


def registered_clusters():
    out = []
    for name in sorted(os.listdir(CSRC)):
        if not re.match(r"cluster_f(32|64)_[a-z]\.cu$", name):
            continue
        for line in open(os.path.join(CSRC, name)):
            line = line.split("//")[0]
            if "push_back" not in line:
                continue
            m = _ENTRY.search(line)
            assert m, f"unparsed registry line in {name}: {line}"
            out.append((m.group(1), m.group(2), int(m.group(4))))  # second mask: every kind the kernel can do
    return out


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_cluster_resident_kernel_runs_on_cpu(tmp_path, oracle):
    cfgs = registered_clusters()
    assert any(name == "float_dsmem_128x128_c4" for _, name, _ in cfgs)  # complex 32768 is NOT this one; 16384 is

    def build_and_run(i):
        inc = tmp_path / f"cl_{i}.inc"
        a, name, kinds = cfgs[i]
        inc.write_text(f'CLUSTER(({a}), "{name}", {kinds}u)\n')
        exe = str(tmp_path / f"cluster_fiber_{i}")
        cmd = ["g++", "-std=c++17", "-O1", "-D__CUDACC__", "-DSSFFT_EMUL", f'-DCLUSTER_CFG_INC="{inc}"',
               "-I" + os.path.join(HOST, "simt"), os.path.join(HOST, "cluster_fiber_emul.cpp"),
               "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe]
        subprocess.run(cmd, check=True, capture_output=True, timeout=900)
        return subprocess.run([exe], capture_output=True, text=True, timeout=1200)

    with ThreadPoolExecutor(min(8, len(cfgs))) as pool:
        results = list(pool.map(build_and_run, range(len(cfgs))))
    runs = 0
    for res in results:
        assert res.returncode == 0 and "CLUSTER-FIBER-EMUL-OK" in res.stdout, res.stdout[-4000:] + res.stderr[-2000:]
        runs += int(res.stdout.split(" runs,")[0].split()[-1])
    assert runs >= 12 * len(cfgs)
