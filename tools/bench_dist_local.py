#!/usr/bin/env python
"""tools/bench_dist_local.py [log2 N] -- BASELINE config 5 through the C ABI of the single-process distributed plan
(ssfft_dist_plan_create / ssfft_dist_exec_c2c): one 2^30-point complex64 transform over 2 / 4 / 8 GPUs of this process.

Per GPU count and chunk count: milliseconds per transform (wall clock around K back-to-back calls + synchronize; the
plan's streams are its own), natural-order and transposed output, and size-independent checks at full size (a single
tone -> N delta, Parseval, ifft(fft(x)) = N x).  One JSON line per configuration; summary in gpurun_out/.
"""
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402
from fft_b200.dist import LocalDistFFT1D  # noqa: E402


def run(n, world, chunks, transposed, iters=5):
    os.environ["SSFFT_DIST_CHUNKS"] = str(chunks)
    per = n // world
    plan = LocalDistFFT1D(n, list(range(world)), transposed_output=transposed)
    xs, ys = [], []
    for r in range(world):
        with torch.cuda.device(r):
            x = torch.empty(per, dtype=torch.complex64, device=f"cuda:{r}")
            fft_b200.fill_uniform(x, 20261017, first_idx=r * per * 2)
            xs.append(x)
            ys.append(torch.empty_like(x))
    for r in range(world):
        torch.cuda.synchronize(r)
    for _ in range(2):
        plan.fft(xs, ys)
    plan.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        plan.fft(xs, ys)
    plan.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / iters
    out = {"n": n, "gpus": world, "chunks": chunks, "transposed_output": transposed, "ms": ms,
           "gflops": 5.0 * n * math.log2(n) / ms / 1e6, "plan": plan.describe()}
    if not transposed:
        # Parseval + round trip on the random input, single tone
        ex = sum((x.real.double() ** 2 + x.imag.double() ** 2).sum().item() for x in xs)
        ey = sum((y.real.double() ** 2 + y.imag.double() ** 2).sum().item() for y in ys)
        backs = [torch.empty_like(x) for x in xs]
        plan.ifft(ys, backs)
        plan.synchronize()
        num = sum(((b.real.double() - n * x.real.double()) ** 2 + (b.imag.double() - n * x.imag.double()) ** 2).sum().item()
                  for b, x in zip(backs, xs))
        del backs
        f0 = 123456789 % n
        tones = []
        for r in range(world):
            idx = torch.arange(per, device=f"cuda:{r}", dtype=torch.int64) + r * per
            ph = ((idx * f0) % n).double() * (2.0 * math.pi / n)
            tones.append(torch.complex(torch.cos(ph), torch.sin(ph)).to(torch.complex64))
            del idx, ph
        touts = [torch.empty_like(t) for t in tones]
        for r in range(world):
            torch.cuda.synchronize(r)  # the plan runs on its own streams: the tones must exist before it reads them
        plan.fft(tones, touts)
        plan.synchronize()
        e_all = sum((t.real.double() ** 2 + t.imag.double() ** 2).sum().item() for t in touts)
        own = f0 // per
        pk = touts[own][f0 - own * per].item()
        tone_err = math.sqrt(max(e_all - abs(pk) ** 2, 0.0) + (pk.real - n) ** 2 + pk.imag ** 2) / n
        lim = 1e-6 * math.log2(n)
        out["checks"] = {"tone_rel_err": tone_err, "parseval_rel_err": abs(ey / (n * ex) - 1.0),
                         "roundtrip_rel_l2": math.sqrt(num / (n * n * ex)), "tolerance": lim}
        out["checks"]["ok"] = bool(tone_err <= lim and out["checks"]["roundtrip_rel_l2"] <= 2 * lim and out["checks"]["parseval_rel_err"] < 1e-4)
    plan.close()
    del xs, ys
    torch.cuda.empty_cache()
    return out


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    n = 1 << lg
    rows = []
    ngpu = torch.cuda.device_count()
    for world in (2, 4, 8):
        if world > ngpu:
            continue
        for chunks in [int(c) for c in os.environ.get("SSFFT_BENCH_DIST_CHUNKS", "1,2,4,8").split(",")]:
            for transposed in (False, True):
                if transposed and chunks not in (1, 4):
                    continue
                try:
                    r = run(n, world, chunks, transposed)
                except Exception as exc:
                    r = {"n": n, "gpus": world, "chunks": chunks, "transposed_output": transposed, "error": repr(exc)[:300]}
                rows.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/bench_dist_local_{lg}_{ngpu}gpu.json", "w"), indent=1)


if __name__ == "__main__":
    main()
