#!/usr/bin/env python
"""tools/prof_one.py KIND N [BATCH] -- launch one transform kind a few times (target for ncu captures).
KIND: c2c | r2c | c2r.  BATCH defaults to ~256 MiB of input.  SSFFT_PROF_PREC=float64 for double precision."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402

kind, n = sys.argv[1], int(sys.argv[2])
batch = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, (1 << 28) // (n * (8 if kind == "c2c" else 4)))
prec = os.environ.get("SSFFT_PROF_PREC", "float32")
if kind == "c2c":
    f = fft_b200.FFT(n, dtype=prec)
    x = torch.empty((batch, n), dtype=torch.complex64 if prec == "float32" else torch.complex128, device="cuda")
    y = torch.empty_like(x)
    fft_b200.fill_uniform(x, 1)
    run = lambda: f.fft(x, y)  # noqa: E731
else:
    f = fft_b200.RealFFT(n)
    x = torch.empty((batch, n), dtype=torch.float32, device="cuda")
    y = torch.empty((batch, n // 2), dtype=torch.complex64, device="cuda")
    fft_b200.fill_uniform(x, 1)
    f.fft(x, y)
    run = (lambda: f.fft(x, y)) if kind == "r2c" else (lambda: f.ifft(y, x))
print(f.describe())
for _ in range(4):
    run()
torch.cuda.synchronize()
