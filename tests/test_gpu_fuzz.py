"""GPU sweep over every size that has a specialised kernel x awkward batch counts (1, 2, odd, just past a CTA's group
size, several persistent waves): forward against the oracle for one transform of each batch, and ifft(fft(x)) == N x
for all of them; RealFFT forward / inverse likewise.  Catches configuration mistakes in the kernel registry (ragged
last groups of the TMA prefetch, transforms-per-CTA edge cases) that the fixed-batch parity tests would miss."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import fft_b200  # noqa: E402

POW2 = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
FAST = [96, 192, 384, 768, 1536, 3072, 6144, 144, 288, 576, 1152, 2304, 4608, 9216]
MIXED = [500, 1000, 2187, 3000, 3125, 6000]
BATCHES = [1, 2, 5, 33, 257, 1500]


def tol(n, single):
    return (1e-6 if single else 1e-14) * max(1.0, math.log2(n))


@pytest.mark.parametrize("prec", ["float32", "float64"])
def test_every_registered_size_and_awkward_batches(oracle, cuda_device, prec):
    single = prec == "float32"
    cdt = torch.complex64 if single else torch.complex128
    rdt = torch.float32 if single else torch.float64
    sizes = POW2 + FAST + MIXED if single else [n for n in POW2 + FAST + MIXED if n <= 8192]
    for n in sizes:
        f = fft_b200.FFT(n, dtype=prec)
        r = fft_b200.RealFFT(2 * n, dtype=prec)  # its complex core is the same length-n kernel, real flavours
        for batch in BATCHES:
            if n * batch > (1 << 24):
                continue
            x = torch.empty((batch, n), dtype=cdt, device="cuda")
            fft_b200.fill_uniform(x, 1000 + batch)
            y = torch.empty_like(x)
            z = torch.empty_like(x)
            f.fft(x, y)
            f.ifft(y, z)
            err = (torch.linalg.vector_norm(z - n * x, dim=1) / torch.linalg.vector_norm(n * x, dim=1)).max().item()
            assert err <= 2 * tol(n, single), (prec, n, batch, "round trip", err, f.describe())
            last = x[batch - 1:].cpu().numpy()
            ref = oracle.run(oracle.KIND_C2C_FWD, last, n, threads=1)[0]
            e1 = oracle.rel_l2(y[batch - 1:].cpu().numpy(), ref)
            assert e1 <= tol(n, single), (prec, n, batch, "forward", e1, f.describe())
            xr = torch.empty((batch, 2 * n), dtype=rdt, device="cuda")
            fft_b200.fill_uniform(xr, 2000 + batch)
            sp = torch.empty((batch, n), dtype=cdt, device="cuda")
            back = torch.empty_like(xr)
            r.fft(xr, sp)
            r.ifft(sp, back)
            err = (torch.linalg.vector_norm(back - 2 * n * xr, dim=1) / torch.linalg.vector_norm(2 * n * xr, dim=1)).max().item()
            assert err <= 2 * tol(2 * n, single), (prec, n, batch, "real round trip", err, r.describe())
            refr = oracle.rfft(xr[batch - 1:].cpu().numpy())
            e2 = oracle.rel_l2(sp[batch - 1:].cpu().numpy(), refr)
            assert e2 <= tol(2 * n, single), (prec, n, batch, "real forward", e2, r.describe())
