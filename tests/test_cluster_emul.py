"""CPU test: the cluster-resident four-step kernel (fft_b200/csrc/cluster.cuh) executed phase by phase on the host.

tests/host/cluster_emul.cu instantiates the kernel's __host__ __device__ phases for every registered geometry,
plays all CTAs x threads of a cluster sequentially and compares with a double-precision DFT (C2C both directions,
RealFFT forward and inverse); it also replays every shared-memory access pattern through a half-warp bank model
and requires zero conflicts.  nvcc compiles it here without a GPU.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_cluster_kernel_index_logic_on_cpu(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "cluster_emul")
    subprocess.run([nvcc, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                    os.path.join(ROOT, "tests", "host", "cluster_emul.cu"), "-o", exe], check=True, capture_output=True,
                   timeout=600)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "CLUSTER-EMUL-OK" in res.stdout, res.stdout + res.stderr
    assert "extra (conflict) wavefronts 0" in res.stdout
