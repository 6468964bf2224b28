"""GPU check of the cluster-resident four-step (fft_b200/csrc/cluster.cuh), run by tests/test_gpu_parity.py in a
fresh process with SSFFT_DSMEM_ALL=1 so that every registered geometry and kind is exercised (by default the
planner only picks the cluster-resident kernel where it measured faster than the alternatives)."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402
from oracle import oracle  # noqa: E402  (checker)

SEED = 7


def tol(n, dtype):
    lg = max(1.0, math.log2(max(n, 2)))
    return (1e-6 if dtype in (np.complex64, np.float32) else 1e-14) * lg


def main():
    rng_idx = lambda b: sorted({0, 1, 2, b // 2, b - 2, b - 1})  # noqa: E731  oracle-checked transforms
    for n in (16384, 32768, 65536):
        batch = 301
        f = fft_b200.FFT(n)
        if n <= 32768:
            assert "cluster-resident" in f.describe(), f.describe()
        x = torch.empty((batch, n), dtype=torch.complex64, device="cuda")
        fft_b200.fill_uniform(x, SEED)
        y = torch.empty_like(x)
        z = torch.empty_like(x)
        f.fft(x, y)
        f.ifft(y, z)
        torch.cuda.synchronize()
        idx = rng_idx(batch)
        xs = x[idx].cpu().numpy()
        ref = oracle.run(oracle.KIND_C2C_FWD, xs, n, threads=6)[0]
        assert oracle.rel_l2(y[idx].cpu().numpy(), ref) <= tol(n, np.complex64), (n, f.describe())
        refi = oracle.run(oracle.KIND_C2C_INV, xs, n, threads=6)[0]
        w = torch.empty_like(x)
        f.ifft(x, w)
        assert oracle.rel_l2(w[idx].cpu().numpy(), refi) <= tol(n, np.complex64), (n, "inverse")
        # whole batch: ifft(fft(x)) == N x, per transform
        err = (torch.linalg.vector_norm(z - n * x, dim=1) / torch.linalg.vector_norm(n * x, dim=1)).max().item()
        assert err <= 2 * tol(n, np.complex64), (n, "round trip", err)
        del x, y, z, w
    for n in (32768, 65536, 131072):
        batch = 203
        r = fft_b200.RealFFT(n)
        if n <= 65536:
            assert "cluster-resident" in r.describe(), r.describe()
        x = torch.empty((batch, n), dtype=torch.float32, device="cuda")
        fft_b200.fill_uniform(x, SEED + 1)
        spec = torch.full((batch * (n // 2) + 64,), 123.0 + 456.0j, dtype=torch.complex64, device="cuda")
        out = spec[: batch * (n // 2)].view(batch, n // 2)
        r.fft(x, out)
        back = torch.empty_like(x)
        r.ifft(out, back)
        torch.cuda.synchronize()
        assert torch.all(spec[batch * (n // 2):] == 123.0 + 456.0j), "wrote past the N/2 bins"
        idx = rng_idx(batch)
        xs = x[idx].cpu().numpy()
        ref = oracle.rfft(xs)
        assert oracle.rel_l2(out[idx].cpu().numpy(), ref) <= tol(n, np.float32), (n, r.describe())
        assert oracle.rel_l2(back[idx].cpu().numpy(), oracle.irfft(ref)) <= 2 * tol(n, np.float32), (n, "inverse")
        err = (torch.linalg.vector_norm(back - n * x, dim=1) / torch.linalg.vector_norm(n * x, dim=1)).max().item()
        assert err <= 2 * tol(n, np.float32), (n, "round trip", err)
        del x, spec, out, back
    print("CLUSTER-GPU-OK")


if __name__ == "__main__":
    main()
