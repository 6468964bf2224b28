#!/usr/bin/env python
"""tools/sanitize_smoke.py -- small transforms through every kernel family, meant to be run under compute-sanitizer
(memcheck / racecheck / synccheck).  Checks results against a float64 torch-free DFT identity (ifft(fft(x)) == N x)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402

cases = [("c2c", 256, 40), ("c2c", 1000, 9), ("c2c", 2048, 13), ("c2c", 4096, 7), ("c2c", 8192, 5), ("c2c", 16384, 3),
         ("c2c", 32768, 10), ("c2c", 65536, 5), ("c2c", 6000, 3), ("real", 1024, 9), ("real", 8192, 5), ("real", 65536, 5),
         ("real", 32768, 3),
         # round 2: ticket-queue kernels (complex, real in place, 3 * 2^k, 9 * 2^k, fp64), composite plan, Bluestein
         ("c2c", 131072, 3), ("c2c", 1048576, 2), ("real", 131072, 3), ("real", 2097152, 1), ("c2c", 12288, 5), ("c2c", 98304, 3),
         ("real", 49152, 3), ("c2c", 18432, 5), ("c2c", 147456, 2), ("real", 73728, 2), ("c2c64", 16384, 3), ("c2c64", 131072, 2),
         ("real64", 65536, 3), ("c2c64", 24576, 3), ("c2c", 2097152, 1), ("c2c", 1179648, 1), ("c2c", 16411, 2)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:] or str(c[1]) in sys.argv[1:]]
for kind, n, batch in cases:
    prec = "float64" if kind.endswith("64") else "float32"
    cdt = torch.complex128 if prec == "float64" else torch.complex64
    rdt = torch.float64 if prec == "float64" else torch.float32
    lim = 1e-12 if prec == "float64" else 1e-5
    if kind.startswith("c2c"):
        f = fft_b200.FFT(n, dtype=prec)
        x = torch.empty((batch, n), dtype=cdt, device="cuda")
        fft_b200.fill_uniform(x, 3)
        y = torch.empty_like(x)
        z = torch.empty_like(x)
        f.fft(x, y)
        f.ifft(y, z)
        torch.cuda.synchronize()
        err = ((z - n * x).norm() / (n * x).norm()).item()
    else:
        f = fft_b200.RealFFT(n, dtype=prec)
        x = torch.empty((batch, n), dtype=rdt, device="cuda")
        fft_b200.fill_uniform(x, 3)
        y = torch.empty((batch, n // 2), dtype=cdt, device="cuda")
        z = torch.empty_like(x)
        f.fft(x, y)
        f.ifft(y, z)
        torch.cuda.synchronize()
        err = ((z - n * x).norm() / (n * x).norm()).item()
    print(f"{kind} {n} x{batch}: round-trip relL2 {err:.2e}  [{f.describe()[:70]}]", flush=True)
    assert err < lim
print("SANITIZE-SMOKE-OK")
