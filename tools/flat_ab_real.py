#!/usr/bin/env python
"""tools/flat_ab_real.py -- A/B timing of the RealFFT flavours of the ticket-queue four-step (flat.cuh).

python tools/flat_ab_real.py <tag> [real sizes...]: per size ~1 GiB of float32 input, forward and inverse timed apart
(CUDA events, 10 launches after 3 warm-ups), one line per registered variant (SSFFT_FLAT_NAME = part of the entry name;
the round-1 cluster kernels with SSFFT_DISABLE_FLAT_REAL=1).  The round trip ifft(fft(x)) / N is checked against x.
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r"""
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch, fft_b200
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
n = int(sys.argv[1])
batch = max(1, (1 << 30) // (n * 4))
x = torch.empty((batch, n), dtype=torch.float32, device="cuda")
fft_b200.fill_uniform(x, 1)
spec = torch.empty((batch, n // 2), dtype=torch.complex64, device="cuda")
back = torch.empty_like(x)
r = fft_b200.RealFFT(n)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
ms_f = timeit(lambda: r.fft(x, spec))
ms_i = timeit(lambda: r.ifft(spec, back))
k = min(batch, 32)
err = float(torch.linalg.vector_norm(back[:k] / n - x[:k]) / torch.linalg.vector_norm(x[:k]))
byt = 2 * batch * n * 4
print(json.dumps({"n": n, "batch": batch, "ms_fwd": ms_f, "ms_inv": ms_i, "frac_fwd": byt / ms_f / 1e6 / PEAK,
                  "frac_inv": byt / ms_i / 1e6 / PEAK, "roundtrip": err, "plan": r.describe()}))
"""


def run(n, env):
    e = dict(os.environ)
    e.update(env)
    res = subprocess.run([sys.executable, "-c", CHILD, str(n)], capture_output=True, text=True, env=e, timeout=300)
    try:
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception:
        return {"error": (res.stdout + res.stderr)[-600:]}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "flatreal"
    sizes = [int(a) for a in sys.argv[2:]] or [65536]
    configs = [("round-1 cluster kernels", {"SSFFT_DISABLE_FLAT_REAL": "1"}),
               ("default", {}),
               ("ring1 3/SM separate exchange (r1c3x)", {"SSFFT_FLAT_NAME": "r1c3x"}),
               ("ring2 3/SM in place (r2c3i)", {"SSFFT_FLAT_NAME": "r2c3i"}),
               ("ring1 3/SM in place (r1c3i)", {"SSFFT_FLAT_NAME": "r1c3i"}),
               ("wide tiles, 512 threads, ring1 2/SM in place", {"SSFFT_FLAT_NAME": "_w_"})]
    rows = []
    for n in sizes:
        for name, env in configs:
            r = run(n, env)
            r["config"] = name
            rows.append(r)
            if "error" in r:
                print(f"N={n:8d} {name:46s} ERROR {r['error']}", flush=True)
                continue
            print(f"N={n:8d} {name:46s} R2C {r['ms_fwd']:7.4f} ms {100 * r['frac_fwd']:5.1f}%   C2R {r['ms_inv']:7.4f} ms {100 * r['frac_inv']:5.1f}%   "
                  f"round trip {r['roundtrip']:.1e}  [{r['plan'][60:150]}]", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/flat_ab_real_{tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
