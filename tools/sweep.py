#!/usr/bin/env python
"""tools/sweep.py -- throughput of the device path over the north-star size range (development/report tool).

For each N: batched C2C forward and RealFFT forward+inverse, ~1 GiB of input per case, CUDA-event timed.
Prints a table and writes gpurun_out/sweep_<tag>.json.  Fractions are algorithmic bytes / time / measured HBM peak.
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402

# A/B measurements of compile-time experiment switches (tools/ab_variants.sh): time another BUILD of the same library.
# The override lives in this development tool, not in the package: the product always loads fft_b200/libssfft.so.
if os.environ.get("SSFFT_LIB"):
    from fft_b200 import _lib as _L
    _L.LIB_PATH = fft_b200.LIB_PATH = os.environ["SSFFT_LIB"]

PEAK = 6528.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    prec = sys.argv[2] if len(sys.argv) > 2 else "float32"
    sizes = [int(a) for a in sys.argv[3:]] or [256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 1 << 17, 1 << 18,
                                                1 << 19, 1 << 20, 1000, 2187, 3125, 6000, 1536, 2304]
    cdt = torch.complex64 if prec == "float32" else torch.complex128
    rdt = torch.float32 if prec == "float32" else torch.float64
    esz = 8 if prec == "float32" else 16
    rows = []
    for n in sizes:
        batch = max(1, (1 << 30) // (n * esz))
        x = torch.empty((batch, n), dtype=cdt, device="cuda")
        fft_b200.fill_uniform(x, 1)
        y = torch.empty_like(x)
        f = fft_b200.FFT(n, dtype=prec)
        ms = timeit(lambda: f.fft(x, y))
        c2c = {"ms": ms, "gflops": batch * 5 * n * math.log2(n) / ms / 1e6, "frac": 2 * batch * n * esz / ms / 1e6 / PEAK,
               "plan": f.describe()}
        del x, y
        row = {"n": n, "batch": batch, "c2c": c2c}
        if n % 2 == 0:
            nr = n
            br = max(1, (1 << 30) // (nr * esz // 2))
            xr = torch.empty((br, nr), dtype=rdt, device="cuda")
            fft_b200.fill_uniform(xr, 2)
            spec = torch.empty((br, nr // 2), dtype=cdt, device="cuda")
            back = torch.empty_like(xr)
            r = fft_b200.RealFFT(nr, dtype=prec)
            ms_f = timeit(lambda: r.fft(xr, spec))
            ms_i = timeit(lambda: r.ifft(spec, back))
            bytes_dir = 2 * br * nr * (esz // 2)
            row["real"] = {"batch": br, "ms_fwd": ms_f, "ms_inv": ms_i, "frac_fwd": bytes_dir / ms_f / 1e6 / PEAK,
                           "frac_inv": bytes_dir / ms_i / 1e6 / PEAK,
                           "gflops_fwd_inv": br * 5 * nr * math.log2(nr) / (ms_f + ms_i) / 1e6, "plan": r.describe()}
            del xr, spec, back
        rows.append(row)
        rr = row.get("real")
        print(f"N={n:8d}  C2C {c2c['ms']:8.4f} ms {c2c['gflops']:9.0f} GF/s {100 * c2c['frac']:5.1f}%"
              + (f"   | R2C {100 * rr['frac_fwd']:5.1f}%  C2R {100 * rr['frac_inv']:5.1f}%" if rr else "")
              + f"   [{c2c['plan'][:60]}]", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"peak_gbs": PEAK, "precision": prec, "rows": rows}, open(f"gpurun_out/sweep_{tag}_{prec}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
