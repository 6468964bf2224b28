#!/usr/bin/env python
"""tools/flat_stats.py -- where the time of the ticket-queue four-step goes (measurement builds of the library).

libssfft_stats.so (-DSSFFT_FLAT_STATS=1) counts, per CTA, the cycles its consumers waited for a ring slot to fill, the
latency of the ticket atomic, of the completion signal, of the dependency polls and of the copies the consumers had to
wait for; libssfft_nocompute.so additionally drops the butterflies (same copies, barriers, stores: what the schedule and
the memory system can do without the arithmetic).  Build both with tools/gpu_r02d.sh's make lines.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from flat_ab import run  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = "SSFFT_FLAT_VARIANT"
VARIANTS = [("ring2 3/SM in place", {}), ("ring1 3/SM separate", {V: "1,3,0"}), ("ring3 2/SM in place", {V: "3,2,1"})]

for lib in ("libssfft_stats.so", "libssfft_nocompute.so", "libssfft_nodeps.so"):
    path = os.path.join(ROOT, "fft_b200", lib)
    if not os.path.exists(path):
        continue
    for n in [int(a) for a in sys.argv[1:]] or [65536, 1 << 18, 1 << 20]:
        for name, env in VARIANTS:
            e = dict(env, SSFFT_LIB=path, SSFFT_FLAT_STATS_PRINT="1")
            r = run(n, e)
            if "error" in r:
                print(f"{lib} N={n} {name}: ERROR {r['error']}", flush=True)
            else:
                print(f"{lib:22s} N={n:8d} {name:22s} {r['ms']:8.4f} ms {100 * r['frac']:5.1f}%", flush=True)
