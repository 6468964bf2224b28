"""CPU tests (no GPU): pin the oracle (oracle/) to the reference.

Three anchors, per the build contract:
  1. golden outputs generated from the unmodified reference header (tests/golden/reference_vectors.npz,
     made by tests/golden/make_golden.py) -- always available, also on the GPU box;
  2. the compiled reference itself (oracle/_ref/libssfft_ref.so) when present: bit-exact comparison;
  3. the known-answer properties of the reference's own tests (tests/00-fft.cpp, tests/01-real.cpp).
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")
SEED = 7  # must match tests/golden/make_golden.py

# tests/00-fft.cpp:8-16
REF_TEST_SIZES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 3, 6, 9, 12, 18, 24, 5, 10, 15, 20, 25, 7, 14, 21, 28, 49,
                  11, 13, 17, 19, 22, 23]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_golden_c2c_bit_exact(oracle, golden):
    """The C restatement reproduces the reference's outputs bit for bit on every stored vector."""
    n_checked = 0
    for key in golden.files:
        if not key.startswith("c2c_"):
            continue
        _, direction, prec, n = key.split("_")
        n = int(n)
        cdt = np.complex64 if prec == "f32" else np.complex128
        x = oracle.uniform_complex((1, n), SEED, cdt)
        y = oracle.run(oracle.KIND_C2C_INV if direction == "inv" else oracle.KIND_C2C_FWD, x, n)[0][0]
        assert np.array_equal(y, golden[key]), key
        n_checked += 1
    assert n_checked >= 100


def test_golden_real_bit_exact(oracle, golden):
    n_checked = 0
    for key in golden.files:
        kind, prec, n = key.split("_")[0], key.split("_")[1], key.split("_")[-1]
        if kind not in ("r2c", "m2c", "c2r", "c2m"):
            continue
        n = int(n)
        rdt = np.float32 if prec == "f32" else np.float64
        x = oracle.uniform(n, SEED, rdt).reshape(1, n)
        modified = kind in ("m2c", "c2m")
        y = oracle.rfft(x, modified)
        if kind in ("r2c", "m2c"):
            assert np.array_equal(y[0], golden[key]), key
        else:
            assert np.array_equal(oracle.irfft(y, modified)[0], golden[key]), key
        n_checked += 1
    assert n_checked >= 400


def test_matches_compiled_reference_when_present(oracle):
    """Direct bit-exact comparison with oracle/_ref (skipped on boxes where it was never built)."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built on this machine")
    for cdt in (np.complex64, np.complex128):
        for n in [1, 2, 3, 5, 17, 30, 64, 100, 243, 1000, 1024, 2187, 3125, 4096, 6000, 32768]:
            x = oracle.uniform_complex((3, n), 11, cdt)
            for kind in (oracle.KIND_C2C_FWD, oracle.KIND_C2C_INV):
                a = oracle.run(kind, x, n, 1, "port")[0]
                b = oracle.run(kind, x, n, 2, "reference")[0]
                assert np.array_equal(a, b), (cdt, n, kind)
    for rdt in (np.float32, np.float64):
        for n in [2, 4, 6, 10, 22, 98, 256, 1000, 65536]:
            x = oracle.uniform(2 * n, 12, rdt).reshape(2, n)
            for mod in (False, True):
                a = oracle.rfft(x, mod)
                assert np.array_equal(a, oracle.rfft(x, mod, "reference")), (rdt, n, mod)
                assert np.array_equal(oracle.irfft(a, mod), oracle.irfft(a, mod, "reference")), (rdt, n, mod)


def test_individual_bins(oracle):
    """tests/00-fft.cpp:19-60 -- e^{+2 pi i n bin/N} -> N delta[bin]; input untouched."""
    for n in REF_TEST_SIZES:
        k = np.arange(n)
        x = np.exp(2j * np.pi * np.outer(k, k) / n).astype(np.complex128)  # row `bin`
        x_copy = x.copy()
        y = oracle.fft(x)
        assert np.array_equal(x, x_copy)
        assert np.abs(y - n * np.eye(n)).max() < 1e-9 * max(n, 1), n


def test_linearity_and_inverse(oracle):
    """tests/00-fft.cpp:62-145 -- F(a+b) = F(a)+F(b); ifft(fft(x)) = N x (unnormalised both ways)."""
    for n in REF_TEST_SIZES:
        a = oracle.uniform_complex((1, n), 1, np.complex128)
        b = oracle.uniform_complex((1, n), 2, np.complex128)
        assert oracle.rel_l2(oracle.fft(a + b), oracle.fft(a) + oracle.fft(b)) < 1e-13
        assert oracle.rel_l2(oracle.ifft(oracle.fft(a)), n * a) < 1e-13


def test_against_numpy(oracle):
    """Independent cross-check of sign and scaling conventions against numpy's FFT."""
    for n in [1, 2, 7, 12, 49, 256, 1000, 2187, 4096]:
        x = oracle.uniform_complex((2, n), 3, np.complex128)
        assert oracle.rel_l2(oracle.fft(x), np.fft.fft(x)) < 1e-13
        assert oracle.rel_l2(oracle.ifft(x), np.fft.ifft(x) * n) < 1e-13
        xf = x.astype(np.complex64)
        assert oracle.rel_l2(oracle.fft(xf), np.fft.fft(xf.astype(np.complex128))) < 1.2e-6


def test_real_matches_complex(oracle):
    """tests/01-real.cpp:17-96 -- bins 1..N/2-1 equal the complex FFT; bin 0 packs (DC, Nyquist)."""
    for n in range(2, 100, 2):
        x = oracle.uniform(n, 4, np.float64).reshape(1, n)
        full = np.fft.fft(x[0])
        y = oracle.rfft(x)[0]
        assert abs(y[0].real - full[0].real) < 1e-12 * n and abs(y[0].imag - full[n // 2].real) < 1e-12 * n
        assert np.abs(y[1:] - full[1:n // 2]).max(initial=0) < 1e-12 * n
        assert oracle.rel_l2(oracle.irfft(y[None]), n * x) < 1e-13
        # modified: spectrum of x[n] * exp(-i pi n / N) (tests/01-real.cpp:42-48)
        rot = np.exp(-1j * np.pi * np.arange(n) / n)
        fullm = np.fft.fft(x[0] * rot)
        ym = oracle.rfft(x, modified=True)[0]
        assert np.abs(ym - fullm[:n // 2]).max() < 1e-12 * n
        assert oracle.rel_l2(oracle.irfft(ym[None], modified=True), n * x) < 1e-13


def test_size_helpers(oracle):
    """tests/00-fft.cpp:168-187 and the RealFFT quirks recorded in SURVEY.md section 8a row R0."""
    lib = oracle.port()

    def fast(v):
        c3 = c5 = 0
        while v % 2 == 0 and v > 1:
            v //= 2
        while v % 3 == 0 and v > 1:
            v //= 3
            c3 += 1
        while v % 5 == 0 and v > 1:
            v //= 5
            c5 += 1
        return v == 1 and c3 + c5 <= 2

    for i in range(1, 1000):
        above, below = lib.oracle_fft_size_minimum(i), lib.oracle_fft_size_maximum(i)
        assert above >= i and below <= i and fast(above) and fast(below)
    assert lib.oracle_fft_size_minimum(1025) == 1152 and lib.oracle_fft_size_maximum(1025) == 1024
    assert lib.oracle_fft_size_minimum(1000) == 1024 and lib.oracle_fft_size_maximum(1000) == 768
    assert lib.oracle_fft_size_minimum(6000) == 6144 and lib.oracle_fft_size_maximum(3125) == 3072
    assert lib.oracle_realfft_size_minimum(256) == 258 and lib.oracle_realfft_size_minimum(1000) == 1026
    assert lib.oracle_realfft_size_maximum(1000) == 1024
    if oracle.have_reference():
        ref = oracle.reference()
        for i in range(1, 5000):
            assert lib.oracle_fft_size_minimum(i) == ref.ref_fft_size_minimum(i)
            assert lib.oracle_fft_size_maximum(i) == ref.ref_fft_size_maximum(i)
            assert lib.oracle_realfft_size_minimum(i) == ref.ref_realfft_size_minimum(i)
            assert lib.oracle_realfft_size_maximum(i) == ref.ref_realfft_size_maximum(i)
        assert ref.ref_realfft_setsize_return(64) == 32 and ref.ref_realfft_size(7) == 6


def test_plan_structure(oracle):
    """factor lists, step lists and the permutation == mixed-radix
    digit reversal (the fact the CUDA kernels rely on to fold the permutation into their indexing)."""
    assert oracle.plan_info(1000)[0] == [2, 2, 2, 5, 5, 5]
    assert oracle.plan_info(2187)[0] == [3] * 7
    assert oracle.plan_info(6000)[0] == [2, 2, 2, 2, 3, 5, 5, 5]
    factors, steps, ntw, _ = oracle.plan_info(4096, "f32")
    assert factors == [2] * 12 and ntw == 5460
    assert [(s[0], s[3], s[4]) for s in steps] == [(4, 1, 1024), (4, 4, 256), (4, 16, 64), (4, 64, 16), (4, 256, 4),
                                                   (4, 1024, 1)]
    _, steps, _, _ = oracle.plan_info(1000)
    assert [(s[0], s[1], s[3]) for s in steps] == [(0, 5, 1), (0, 5, 5), (0, 5, 25), (2, 2, 125), (4, 4, 250)]
    assert len(oracle.plan_info(65536, "f32")[1]) == 29  # cache-blocking branch (:130-133)
    for n in [12, 30, 64, 100, 1000, 2187]:
        factors, _, _, perm = oracle.plan_info(n)
        # source index n = d0 + f0*d1 + f0*f1*d2 + ...  ->  destination d0*N/f0 + d1*N/(f0 f1) + ...
        src = np.arange(n)
        dest = np.zeros(n, dtype=np.int64)
        rem, scale = src.copy(), n
        for f in factors:
            scale //= f
            dest += (rem % f) * scale
            rem //= f
        assert np.array_equal(perm, dest), n


def test_generator_twins(oracle):
    lib = oracle.port()
    for dt, fn in ((np.float32, lib.oracle_fill_uniform_f32), (np.float64, lib.oracle_fill_uniform_f64)):
        a = oracle.uniform(4097, 20261017, dt, first_idx=12345)
        b = np.empty(4097, dt)
        fn(b.ctypes.data, 4097, 20261017, 12345)
        assert np.array_equal(a, b)
        assert a.min() >= -0.5 and a.max() < 0.5 and abs(a.mean()) < 0.02


def test_batch_threads_agree(oracle):
    x = oracle.uniform_complex((13, 360), 5, np.complex64)
    a = oracle.run(oracle.KIND_C2C_FWD, x, 360, 1)[0]
    b = oracle.run(oracle.KIND_C2C_FWD, x, 360, 4)[0]
    assert np.array_equal(a, b)
    assert oracle.run(oracle.KIND_C2C_FWD, x[:0], 360, 4)[0].shape == (0, 360)
