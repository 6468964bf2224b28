"""Throughput of the extended execution calls (ssfft_exec_*_ex) against their unfused equivalents.

    python tools/bench_ex.py [tag [stft]]  -> gpurun_out/bench_ex_<tag>.json  (one JSON object per line)

Workloads (float32, synthetic uniform data, buffers larger than L2):
  stft     RealFFT N=1024, hop 256, Hann-like window: frames read straight out of the signal (fused) vs
           SSFFT_EX_UNFUSED=1 (gather pass + plain transform) vs the plain transform on frames that already exist
  conv     RealFFT N=4096: ifft(spectrum * filter) * window, filter per transform, fused vs unfused
  columns  FFT N=4096 down the columns of a row-major [4096, cols] matrix, in place, fused vs unfused
Algorithmic bytes = what has to cross HBM once: every distinct input element, every multiplier table that is not
shared, every output element.  Timed with CUDA events on the launching stream, 3 warm-ups, median of 10.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fft_b200  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6528.1


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


def with_unfused(fn):
    os.environ["SSFFT_EX_UNFUSED"] = "1"
    try:
        return timed(fn)
    finally:
        del os.environ["SSFFT_EX_UNFUSED"]


def line(out, name, variant, ms, alg_bytes, launches, extra=None):
    rec = {"workload": name, "variant": variant, "ms": round(ms, 4), "algorithmic_GB": round(alg_bytes / 1e9, 3),
           "achieved_GBps": round(alg_bytes / ms / 1e6, 1), "frac_of_hbm_peak": round(alg_bytes / ms / 1e6 / peak_gbs(), 3),
           "launches_per_call": launches}
    rec.update(extra or {})
    out.write(json.dumps(rec) + "\n")
    print(rec)


def launches_of(fn):
    l0 = fft_b200.launch_count()
    fn()
    torch.cuda.synchronize()
    return fft_b200.launch_count() - l0


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    only = sys.argv[2] if len(sys.argv) > 2 else ""  # e.g. "stft": that workload alone (for an ncu capture)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    dev = torch.device("cuda:0")
    with open(os.path.join(ROOT, "gpurun_out", f"bench_ex_{tag}.json"), "w") as out:
        # ---- STFT
        n, hop, frames = 1024, 256, 1 << 19
        h = n // 2
        sig = torch.empty((frames - 1) * hop + n, dtype=torch.float32, device=dev)
        fft_b200.fill_uniform(sig, 20261017)
        win = torch.hann_window(n, dtype=torch.float32, device=dev) + 0.5
        spec = torch.empty((frames, h), dtype=torch.complex64, device=dev)
        rfft = fft_b200.RealFFT(n)
        alg = sig.numel() * 4 + spec.numel() * 8
        desc = {"n": n, "hop": hop, "frames": frames, "plan": rfft.describe()}
        f = lambda: rfft.stft(sig, hop, win, spec)  # noqa: E731
        line(out, "stft", "fused (1 launch: framing + window + R2C)", timed(f), alg, launches_of(f), desc)
        os.environ["SSFFT_EX_UNFUSED"] = "1"
        lu = launches_of(f)
        del os.environ["SSFFT_EX_UNFUSED"]
        line(out, "stft", "unfused (gather pass + plain R2C)", with_unfused(f), alg, lu)
        framed = torch.empty((frames, n), dtype=torch.float32, device=dev)
        fft_b200.fill_uniform(framed, 1)
        g = lambda: rfft.fft(framed, spec)  # noqa: E731
        line(out, "stft", "plain R2C on frames that already exist (no framing, no window)", timed(g), framed.numel() * 4 + spec.numel() * 8,
             launches_of(g))
        del framed, sig, spec
        if only == "stft":
            return
        # ---- fast convolution: ifft(spectrum * filter) * window
        n, batch = 4096, 1 << 17
        h = n // 2
        rfft = fft_b200.RealFFT(n)
        spec = torch.empty((batch, h), dtype=torch.complex64, device=dev)
        filt = torch.empty((batch, h), dtype=torch.complex64, device=dev)
        fft_b200.fill_uniform(spec, 2)
        fft_b200.fill_uniform(filt, 3)
        win = torch.hann_window(n, dtype=torch.float32, device=dev) + 0.5
        y = torch.empty((batch, n), dtype=torch.float32, device=dev)
        alg = spec.numel() * 8 + filt.numel() * 8 + y.numel() * 4
        f = lambda: rfft.ifft_ex(spec.reshape(-1), y.reshape(-1), batch, pre=filt, pre_dist=h, post=win)  # noqa: E731
        line(out, "conv", "fused (1 launch: filter + C2R + window)", timed(f), alg, launches_of(f), {"n": n, "batch": batch, "plan": rfft.describe()})
        line(out, "conv", "unfused (multiply pass + plain C2R + window pass)", with_unfused(f), alg, 3)
        del spec, filt, y
        # ---- column pass of a 2-D transform, in place
        rows, cols = 4096, 1 << 15
        fft = fft_b200.FFT(rows)
        m = torch.empty((rows, cols), dtype=torch.complex64, device=dev)
        fft_b200.fill_uniform(m, 4)
        alg = 2 * m.numel() * 8
        f = lambda: fft.fft_ex(m.reshape(-1), m.reshape(-1), cols, in_stride=cols, in_dist=1, out_stride=cols, out_dist=1)  # noqa: E731
        line(out, "columns", "fused (strided loads / stores inside the kernel)", timed(f), alg, launches_of(f), {"rows": rows, "cols": cols, "plan": fft.describe()})
        line(out, "columns", "unfused (transpose + plain C2C + transpose)", with_unfused(f), alg, 3)
        os.environ["SSFFT_EX_COLCFG"] = "1"  # opt-in experiment: 4 transforms per CTA so that column walks move whole sectors
        try:
            line(out, "columns", "fused, column configuration (SSFFT_EX_COLCFG=1)", timed(f), alg, launches_of(f))
        finally:
            del os.environ["SSFFT_EX_COLCFG"]


if __name__ == "__main__":
    main()
