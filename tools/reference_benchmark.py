#!/usr/bin/env python
"""tools/reference_benchmark.py -- results in the REFERENCE benchmark's own format.

The reference's harness (benchmark/benchmark.h:27-100) times FFT<double>, complex, out-of-place, forward, one
transform per call, over sizes 2^k * {1, 3, 9} <= 2^24 and writes results/<tag>.csv ("size,ops/sec,<name>" + a
column of rate * max(1, N ln N) * 1e-6) and results/<tag>.js.  This tool writes the same files for this library so
the curves overlay the reference's comparison.svg:

    b200-single   one transform per call on device-resident data (what the reference harness measures: latency)
    b200-batched  a batch of transforms per call, rate counted per transform (what the GPU is for: throughput)

    python tools/reference_benchmark.py [outdir] [max_log2] [float64|float32]

float32 writes the "fp32 twin" (b200-single-f32 / b200-batched-f32); both also write <tag>-roofline.txt: the batched rate as
a fraction of the measured HBM peak (2 N sizeof(complex<V>) bytes per transform).
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402


def sizes(max_size):
    out, s = [], 1
    while s <= max_size:
        out.append(s)
        for m in (3, 9):  # benchmark.h:44
            if s * m < max_size:
                out.append(s * m)
        s *= 2
    return sorted(out)


def rate(fn, min_ms=30.0):
    fn()
    torch.cuda.synchronize()
    reps = 4
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if ms >= min_ms or reps >= 1 << 16:
            return reps / (ms * 1e-3)
        reps *= 4


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/results"
    max_size = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
    prec = sys.argv[3] if len(sys.argv) > 3 else "float64"
    suffix = "" if prec == "float64" else "-f32"
    cdt = torch.complex128 if prec == "float64" else torch.complex64
    esz = 16 if prec == "float64" else 8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6439.5
    os.makedirs(outdir, exist_ok=True)
    series = {"b200-single" + suffix: [], "b200-batched" + suffix: []}
    roof = []
    for n in sizes(max_size):
        try:
            plan = fft_b200.FFT(n, dtype=prec)
        except fft_b200.SsfftError as e:
            print(f"N={n}: {e}")
            continue
        batch = max(1, (1 << 29) // (n * esz))
        x = torch.empty((batch, n), dtype=cdt, device="cuda")
        fft_b200.fill_uniform(x, 1)
        y = torch.empty_like(x)
        r1 = rate(lambda: plan.fft(x[:1], y[:1]))
        rb = rate(lambda: plan.fft(x, y)) * batch
        series["b200-single" + suffix].append((n, r1))
        series["b200-batched" + suffix].append((n, rb))
        frac = rb * 2 * n * esz / 1e9 / peak
        roof.append((n, rb, frac, plan.describe()[:90]))
        print(f"N={n:9d}  single {r1:12.1f}/s   batched {rb:14.1f}/s  {100 * frac:5.1f} % of the HBM roofline   [{plan.describe()[:70]}]", flush=True)
        del x, y, plan
    for tag, rows in series.items():
        name = "B200 " + tag.split("-")[1] + (" fp32" if suffix else "")
        with open(os.path.join(outdir, tag + ".csv"), "w") as f:
            f.write(f"size,ops/sec,{name}\n")
            for n, r in rows:
                f.write(f"{n},{r:.15g},{r * max(1.0, n * math.log(n)) * 1e-6:.15g}\n")
        with open(os.path.join(outdir, tag + ".js"), "w") as f:
            f.write(f'addResults("{name}", [')
            f.write(",".join(f"\n\t{{size: {n}, rate: {r * max(1.0, n * math.log(n)) * 1e-6:.6g}}}" for n, r in rows))
            f.write("\n]);")
    with open(os.path.join(outdir, "b200-batched" + suffix + "-roofline.txt"), "w") as f:
        f.write(f"# FFT<{prec}> complex, batched, device-resident; fraction of the measured HBM peak {peak} GB/s (2 N {esz} bytes per transform)\n")
        for n, rb, frac, d in roof:
            f.write(f"{n:9d}  {rb:16.1f} transforms/s  {100 * frac:5.1f} %   {d}\n")


if __name__ == "__main__":
    main()
