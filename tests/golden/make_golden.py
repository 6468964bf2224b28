"""Generate tests/golden/reference_vectors.npz from the UNMODIFIED reference header.

Run where /root/reference exists (this container):   python tests/golden/make_golden.py
It drives oracle/_ref/libssfft_ref.so (oracle/ref_shim.cpp compiled against
/root/reference/signalsmith-fft.h) on inputs from the shared counter-based generator, so only the
OUTPUTS need storing: inputs are regenerated from (seed, dtype, n) by oracle.oracle.uniform*.
The reference holds no golden vectors of its own (its tests draw from rand()); these fixtures pin the
oracle port, and through it the CUDA path, to outputs of the reference itself.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

# the reference's own test sizes (tests/00-fft.cpp:8-16) + BASELINE config sizes + a cache-blocked one
C2C_SIZES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 3, 6, 9, 12, 18, 24, 5, 10, 15, 20, 25, 7, 14, 21, 28, 49,
             11, 13, 17, 19, 22, 23, 1000, 1024, 2187, 3125, 4096, 6000]
REAL_SIZES = list(range(2, 100, 2)) + [256, 1000, 4096]   # tests/01-real.cpp:98-104 uses even 2..98
SEED = 7


def main():
    O.build()
    assert O.have_reference(), "needs /root/reference to build oracle/_ref"
    out = {}
    for prec, cdt, rdt in (("f32", np.complex64, np.float32), ("f64", np.complex128, np.float64)):
        for n in C2C_SIZES:
            if prec == "f64" and n > 4096:
                continue
            x = O.uniform_complex((1, n), SEED, cdt)
            out[f"c2c_fwd_{prec}_{n}"] = O.run(O.KIND_C2C_FWD, x, n, 1, "reference")[0][0]
            out[f"c2c_inv_{prec}_{n}"] = O.run(O.KIND_C2C_INV, x, n, 1, "reference")[0][0]
        for n in REAL_SIZES:
            x = O.uniform(n, SEED, rdt).reshape(1, n)
            for tag, mod in (("r", False), ("m", True)):
                y = O.rfft(x, mod, "reference")
                out[f"{tag}2c_{prec}_{n}"] = y[0]
                out[f"c2{tag}_{prec}_{n}"] = O.irfft(y, mod, "reference")[0]
    # one large float case that exercises the reference's cache-blocking branch (:130-133): N = 65536
    x = O.uniform_complex((1, 65536), SEED, np.complex64)
    out["c2c_fwd_f32_65536"] = O.run(O.KIND_C2C_FWD, x, 65536, 1, "reference")[0][0]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} vectors, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
