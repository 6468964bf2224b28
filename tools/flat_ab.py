#!/usr/bin/env python
"""tools/flat_ab.py -- A/B timing of the ticket-queue four-step (flat.cuh) against the round-1 kernels.

python tools/flat_ab.py <tag> [sizes...]: per size, ~1 GiB of complex64 input, CUDA-event timed (10 launches after 3
warm-ups), one line per configuration: the round-1 path (SSFFT_DISABLE_FLAT=1), then every registered variant
(SSFFT_FLAT_VARIANT = ring,ctas/SM) with the default schedule, then schedule variations of the default variant
(SSFFT_FLAT_DELAY / SSFFT_FLAT_SLOTS / SSFFT_DISCARD).  A result that does not match the round-1 path's is reported.
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r"""
import json, math, os, sys
sys.path.insert(0, os.getcwd())
import torch, fft_b200
if os.environ.get("SSFFT_LIB"):
    from fft_b200 import _lib as _L
    _L.LIB_PATH = fft_b200.LIB_PATH = os.environ["SSFFT_LIB"]
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
n = int(sys.argv[1])
batch = max(1, (1 << 30) // (n * 8))
x = torch.empty((batch, n), dtype=torch.complex64, device="cuda")
fft_b200.fill_uniform(x, 1)
y = torch.empty_like(x)
f = fft_b200.FFT(n)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
ms = timeit(lambda: f.fft(x, y))
chk = float(torch.linalg.vector_norm(y[: min(batch, 64)]).item())
print(json.dumps({"n": n, "batch": batch, "ms": ms, "frac": 2 * batch * n * 8 / ms / 1e6 / PEAK, "chk": chk, "plan": f.describe()}))
"""


def run(n, env):
    e = dict(os.environ)
    e.update(env)
    res = subprocess.run([sys.executable, "-c", CHILD, str(n)], capture_output=True, text=True, env=e, timeout=300)
    for line in res.stderr.splitlines():
        if line.startswith("flat stats"):
            print("      " + line, flush=True)
            break
    try:
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception:
        return {"error": (res.stdout + res.stderr)[-600:]}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "flat"
    sizes = [int(a) for a in sys.argv[2:]] or [32768, 65536]
    V = "SSFFT_FLAT_VARIANT"
    configs = [("round-1 path", {"SSFFT_DISABLE_FLAT": "1"}),
               ("flat ring2 3/SM in place (default)", {}),
               ("flat WIDE tiles (512 threads) ring1 2/SM in place", {"SSFFT_FLAT_NAME": "_w_"}),
               ("flat ring3 2/SM in place", {V: "3,2,1"}),
               ("flat ring1 3/SM separate", {V: "1,3,0"}),
               ("flat ring2 2/SM separate", {V: "2,2,0"}),
               ("default, no discard", {"SSFFT_DISCARD": "0"}),
               ("default, half the delay", {"SSFFT_FLAT_DELAY_PCT": "50"}),
               ("default, twice the delay", {"SSFFT_FLAT_DELAY_PCT": "200"})]
    rows = []
    for n in sizes:
        base = None
        for name, env in configs:
            r = run(n, env)
            r["config"] = name
            rows.append(r)
            if "error" in r:
                print(f"N={n:8d} {name:50s} ERROR {r['error']}", flush=True)
                continue
            if base is None:
                base = r["chk"]
            ok = abs(r["chk"] - base) <= 1e-4 * abs(base)
            print(f"N={n:8d} {name:50s} {r['ms']:8.4f} ms {100 * r['frac']:5.1f}%  {'' if ok else 'RESULT DIFFERS '}[{r['plan'][:70]}]", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/flat_ab_{tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
