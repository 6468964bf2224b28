"""GPU parity tests (-m gpu) of the ticket-queue four-step kernels (fft_b200/csrc/flat.cuh) through the C ABI.

Every registered variant (ring depth x CTAs per SM, SSFFT_FLAT_VARIANT) against the oracle in the same precision, bar
1e-6 * log2(N) relative L2 per transform: batches from one transform (fewer tiles than CTAs) to several times the number
of scratch slots (slots are reused, every CTA loops), forward and inverse, in place, repeated calls on one plan (the
dependency counters are reset per launch), extreme scheduling parameters (no delay / one slot more than the delay), and
-- at a batch of 4096 -- ifft(fft(x)) = N x plus an oracle-checked subset.
"""
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import fft_b200  # noqa: E402

SIZES = [32768, 65536, 2 ** 17, 2 ** 18, 2 ** 19, 2 ** 20]
LONG_SIZES = [2 ** 21, 2 ** 22, 3 * 2 ** 19, 3 * 2 ** 20, 9 * 2 ** 17, 9 * 2 ** 18]  # 2048-point leg (flat_f32_j.cu)
VARIANTS = ["1,3,0", "2,3,1"]  # ring, CTAs per SM, in-place exchange


def tol(n):
    return 1e-6 * math.log2(n)


@pytest.fixture
def flat_env():
    keys = ("SSFFT_FLAT_VARIANT", "SSFFT_FLAT_DELAY", "SSFFT_FLAT_SLOTS", "SSFFT_DISCARD")
    saved = {k: os.environ.get(k) for k in keys}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n", SIZES)
def test_flat_vs_oracle(oracle, cuda_device, flat_env, n, variant):
    flat_env["SSFFT_FLAT_VARIANT"] = variant
    f = fft_b200.FFT(n)
    assert "ticket-queue" in f.describe(), f.describe()
    for batch in ((1, 2, 37, 300) if n <= 65536 else (1, 3, 2 ** 24 // n)):
        x = oracle.uniform_complex((batch, n), 11 + batch, np.complex64)
        xd = torch.from_numpy(x).cuda()
        out = torch.empty_like(xd)
        for inverse in (False, True):
            out.zero_()
            (f.ifft if inverse else f.fft)(xd, out)
            torch.cuda.synchronize()
            assert torch.equal(xd.cpu(), torch.from_numpy(x)), "input was changed"
            ref = oracle.run(oracle.KIND_C2C_INV if inverse else oracle.KIND_C2C_FWD, x, n, threads=8)[0]
            err = oracle.rel_l2(out.cpu().numpy(), ref)
            assert err <= tol(n), (n, variant, batch, inverse, err)
    # same plan again, in place
    x = oracle.uniform_complex((9, n), 5, np.complex64)
    xd = torch.from_numpy(x).cuda()
    f.fft(xd, xd)
    torch.cuda.synchronize()
    assert oracle.rel_l2(xd.cpu().numpy(), oracle.run(oracle.KIND_C2C_FWD, x, n, threads=8)[0]) <= tol(n)


@pytest.mark.parametrize("n", [32768, 65536, 2 ** 18])
def test_flat_scheduling_extremes(oracle, cuda_device, flat_env, n):
    x = oracle.uniform_complex((150 * 65536 // max(n, 65536), n), 3, np.complex64)
    ref = oracle.run(oracle.KIND_C2C_FWD, x, n, threads=8)[0]
    xd = torch.from_numpy(x).cuda()
    for delay, slots, discard in (("0", "1", "1"), ("3", "4", "1"), ("1", "64", "0"), ("40", "41", "1"), ("200", "201", "1")):
        flat_env["SSFFT_FLAT_DELAY"], flat_env["SSFFT_FLAT_SLOTS"], flat_env["SSFFT_DISCARD"] = delay, slots, discard
        f = fft_b200.FFT(n)
        out = torch.zeros_like(xd)
        f.fft(xd, out)
        torch.cuda.synchronize()
        err = oracle.rel_l2(out.cpu().numpy(), ref)
        assert err <= tol(n), (n, delay, slots, err)


@pytest.mark.parametrize("n", [65536, 12288, 24576, 18432])
def test_flat_falls_back_for_unaligned_input(oracle, cuda_device, n):
    """A tensor map needs a 16-byte aligned input: other inputs run the plan's other path (round-1 kernels for powers of
    two, the single-pass interpreter or the composite plan for 3 * 2^k / 9 * 2^k) -- complex and real."""
    batch = 5
    x = oracle.uniform_complex((batch, n), 9, np.complex64)
    buf = torch.zeros(batch * n + 1, dtype=torch.complex64, device="cuda")
    xin = buf[1:].view(batch, n)
    xin.copy_(torch.from_numpy(x))
    assert xin.data_ptr() % 16 == 8
    out = torch.empty((batch, n), dtype=torch.complex64, device="cuda")
    f = fft_b200.FFT(n)
    assert "ticket-queue" in f.describe()
    f.fft(xin, out)
    torch.cuda.synchronize()
    assert oracle.rel_l2(out.cpu().numpy(), oracle.run(oracle.KIND_C2C_FWD, x, n, threads=4)[0]) <= tol(n)
    nr = 2 * n
    xr = oracle.uniform(batch * nr, 10, np.float32).reshape(batch, nr)
    rbuf = torch.zeros(batch * nr + 2, dtype=torch.float32, device="cuda")
    rin = rbuf[2:].view(batch, nr)
    rin.copy_(torch.from_numpy(xr))
    assert rin.data_ptr() % 16 == 8
    spec = torch.empty((batch, n), dtype=torch.complex64, device="cuda")
    r = fft_b200.RealFFT(nr)
    with pytest.raises(fft_b200.SsfftError):  # samples move as pairs: a real buffer must be aligned to one complex value
        r.fft(rbuf[1:batch * nr + 1].view(batch, nr), spec)
    r.fft(rin, spec)
    torch.cuda.synchronize()
    assert oracle.rel_l2(spec.cpu().numpy(), oracle.run(oracle.KIND_R2C, xr, nr, threads=4)[0]) <= tol(nr), r.describe()
    sbuf = torch.zeros(batch * n + 1, dtype=torch.complex64, device="cuda")
    sin = sbuf[1:].view(batch, n)
    sin.copy_(spec)
    back = torch.empty((batch, nr), dtype=torch.float32, device="cuda")
    r.ifft(sin, back)
    torch.cuda.synchronize()
    assert oracle.rel_l2(back.cpu().numpy() / nr, xr) <= 2 * tol(nr), r.describe()


@pytest.mark.parametrize("n", SIZES)
def test_flat_large_batch_properties(oracle, cuda_device, n):
    batch = max(8, 4096 * 65536 // n // 2)
    xd = torch.empty((batch, n), dtype=torch.complex64, device="cuda")
    fft_b200.fill_uniform(xd, 77)
    f = fft_b200.FFT(n)
    y = torch.empty_like(xd)
    z = torch.empty_like(xd)
    f.fft(xd, y)
    f.ifft(y, z)
    torch.cuda.synchronize()
    err = (torch.linalg.vector_norm(z / n - xd) / torch.linalg.vector_norm(xd)).item()
    assert err <= 2 * tol(n), err
    # Parseval per transform
    e_in = torch.linalg.vector_norm(xd, dim=1) ** 2 * n
    e_out = torch.linalg.vector_norm(y, dim=1) ** 2
    assert torch.max(torch.abs(e_out / e_in - 1)).item() < 1e-4
    idx = [0, 1, batch // 2, batch - 2, batch - 1]
    x_sub = xd[idx].cpu().numpy()
    ref = oracle.run(oracle.KIND_C2C_FWD, x_sub, n, threads=5)[0]
    assert oracle.rel_l2(y[idx].cpu().numpy(), ref) <= tol(n)


@pytest.mark.parametrize("n", [2 ** 16, 2 ** 17, 2 ** 18, 2 ** 19, 2 ** 20, 2 ** 21])
def test_flat_real_vs_oracle(oracle, cuda_device, n):
    """RealFFT of length N on the ticket-queue kernels of length N/2: sample pairs through the column tiles, the
    post-twiddle of RealFFT::fft (signalsmith-fft.h:459-472) fused into the row tiles (rows k1 and N1 - k1 side by side in
    the scratch), the pre-twiddle of RealFFT::ifft (:478-492) fused into the column tiles (columns n2 and N2 - n2)."""
    r = fft_b200.RealFFT(n)
    assert "ticket-queue" in r.describe(), r.describe()
    for batch in (1, 3, max(4, 2 ** 23 // n)):
        x = oracle.uniform(batch * n, 13 + batch, np.float32).reshape(batch, n)
        xd = torch.from_numpy(x).cuda()
        spec = torch.zeros((batch, n // 2), dtype=torch.complex64, device="cuda")
        guard = torch.full((batch, n), 7.0, dtype=torch.float32, device="cuda")
        r.fft(xd, spec)
        r.ifft(spec, guard)
        torch.cuda.synchronize()
        assert torch.equal(xd.cpu(), torch.from_numpy(x)), "input was changed"
        ref = oracle.rfft(x)
        err = oracle.rel_l2(spec.cpu().numpy(), ref)
        assert err <= tol(n), (n, batch, err)
        back_ref = oracle.run(oracle.KIND_C2R, ref, n, threads=8)[0]
        err_i = oracle.rel_l2(guard.cpu().numpy(), back_ref)
        assert err_i <= 2 * tol(n), (n, batch, err_i)


def test_flat_real_large_batch_round_trip(oracle, cuda_device):
    n, batch = 65536, 4096
    r = fft_b200.RealFFT(n)
    xd = torch.empty((batch, n), dtype=torch.float32, device="cuda")
    fft_b200.fill_uniform(xd, 3)
    spec = torch.empty((batch, n // 2), dtype=torch.complex64, device="cuda")
    back = torch.empty_like(xd)
    r.fft(xd, spec)
    r.ifft(spec, back)
    torch.cuda.synchronize()
    err = (torch.linalg.vector_norm(back / n - xd) / torch.linalg.vector_norm(xd)).item()
    assert err <= 2 * tol(n), err
    idx = [0, 1, batch // 2, batch - 1]
    assert oracle.rel_l2(spec[idx].cpu().numpy(), oracle.rfft(xd[idx].cpu().numpy())) <= tol(n)


@pytest.mark.parametrize("n", [12288, 24576, 49152, 98304, 196608, 393216, 786432, 18432, 36864, 73728, 147456, 294912, 589824])
def test_flat_three_times_power_of_two(oracle, cuda_device, n):
    """Row stages of length 3 * 2^j (flat_f32_h.cu) and, for 9 * 2^k, column stages with a factor 3 in their last pass
    (flat_f32_i.cu): the 2^k * {3, 9} part of the reference's benchmark sizes
    (benchmark/benchmark.h:27-52) above the single-pass kernels, forward and inverse, several batch shapes, in place, and
    the real transform of twice the length (complex core on these kernels)."""
    f = fft_b200.FFT(n)
    assert "ticket-queue" in f.describe(), f.describe()
    for batch in (1, 5, max(2, 2 ** 23 // n)):
        x = oracle.uniform_complex((batch, n), 41 + batch, np.complex64)
        xd = torch.from_numpy(x).cuda()
        out = torch.empty_like(xd)
        for inverse in (False, True):
            out.zero_()
            (f.ifft if inverse else f.fft)(xd, out)
            torch.cuda.synchronize()
            assert torch.equal(xd.cpu(), torch.from_numpy(x)), "input was changed"
            ref = oracle.run(oracle.KIND_C2C_INV if inverse else oracle.KIND_C2C_FWD, x, n, threads=8)[0]
            err = oracle.rel_l2(out.cpu().numpy(), ref)
            assert err <= tol(n), (n, batch, inverse, err, f.describe())
    x = oracle.uniform_complex((7, n), 6, np.complex64)
    xd = torch.from_numpy(x).cuda()
    launches = fft_b200.launch_count()
    f.fft(xd, xd)
    torch.cuda.synchronize()
    assert fft_b200.launch_count() - launches == 1, "the plan's fallback path ran instead of the one persistent launch"
    assert oracle.rel_l2(xd.cpu().numpy(), oracle.run(oracle.KIND_C2C_FWD, x, n, threads=8)[0]) <= tol(n)
    r = fft_b200.RealFFT(2 * n)
    xr = oracle.uniform(3 * 2 * n, 8, np.float32).reshape(3, 2 * n)
    xrd = torch.from_numpy(xr).cuda()
    spec = torch.empty((3, n), dtype=torch.complex64, device="cuda")
    back = torch.empty_like(xrd)
    r.fft(xrd, spec)
    r.ifft(spec, back)
    torch.cuda.synchronize()
    assert oracle.rel_l2(spec.cpu().numpy(), oracle.run(oracle.KIND_R2C, xr, 2 * n, threads=8)[0]) <= tol(2 * n), r.describe()
    assert oracle.rel_l2(back.cpu().numpy() / (2 * n), xr) <= 2 * tol(2 * n)


@pytest.mark.parametrize("n", [2 ** 14, 2 ** 15, 2 ** 16, 2 ** 17, 2 ** 18, 2 ** 19, 2 ** 20, 12288, 24576, 49152, 18432, 36864, 98304, 196608])
def test_flat_double_precision(oracle, cuda_device, n):
    """fp64 on the ticket-queue kernels (flat_f64_a.cu: TMA boxes of two 8-byte words per element): complex forward /
    inverse / in place and the real transform of twice the length, against the oracle at 1e-14 * log2 N."""
    lim = 1e-14 * math.log2(n)
    f = fft_b200.FFT(n, dtype="float64")
    assert "ticket-queue" in f.describe(), f.describe()
    for batch in (1, 3, max(2, 2 ** 22 // n)):
        x = oracle.uniform_complex((batch, n), 51 + batch, np.complex128)
        xd = torch.from_numpy(x).cuda()
        out = torch.empty_like(xd)
        for inverse in (False, True):
            out.zero_()
            launches = fft_b200.launch_count()
            (f.ifft if inverse else f.fft)(xd, out)
            torch.cuda.synchronize()
            assert fft_b200.launch_count() - launches == 1
            assert torch.equal(xd.cpu(), torch.from_numpy(x)), "input was changed"
            ref = oracle.run(oracle.KIND_C2C_INV if inverse else oracle.KIND_C2C_FWD, x, n, threads=8)[0]
            err = oracle.rel_l2(out.cpu().numpy(), ref)
            assert err <= lim, (n, batch, inverse, err, f.describe())
    x = oracle.uniform_complex((5, n), 7, np.complex128)
    xd = torch.from_numpy(x).cuda()
    f.fft(xd, xd)
    torch.cuda.synchronize()
    assert oracle.rel_l2(xd.cpu().numpy(), oracle.run(oracle.KIND_C2C_FWD, x, n, threads=8)[0]) <= lim
    r = fft_b200.RealFFT(2 * n, dtype="float64")
    if n & (n - 1) == 0:  # the fp64 entries of 3 * 2^k / 9 * 2^k carry no real kernels: their real transforms wrap the complex one
        assert "ticket-queue" in r.describe(), r.describe()
    xr = oracle.uniform(3 * 2 * n, 9, np.float64).reshape(3, 2 * n)
    xrd = torch.from_numpy(xr).cuda()
    spec = torch.empty((3, n), dtype=torch.complex128, device="cuda")
    back = torch.empty_like(xrd)
    r.fft(xrd, spec)
    r.ifft(spec, back)
    torch.cuda.synchronize()
    assert oracle.rel_l2(spec.cpu().numpy(), oracle.run(oracle.KIND_R2C, xr, 2 * n, threads=8)[0]) <= 1e-14 * math.log2(2 * n), r.describe()
    assert oracle.rel_l2(back.cpu().numpy() / (2 * n), xr) <= 2e-14 * math.log2(2 * n)


@pytest.mark.parametrize("n", LONG_SIZES)
def test_flat_2048_point_leg(oracle, cuda_device, n):
    """2^21, 2^22, 3 * 2^19 on the ticket-queue kernels: forward / inverse, batch 1 and 5, in place."""
    f = fft_b200.FFT(n)
    assert "ticket-queue" in f.describe(), f.describe()
    for batch in (1, 5):
        x = oracle.uniform_complex((batch, n), 71 + batch, np.complex64)
        xd = torch.from_numpy(x).cuda()
        out = torch.empty_like(xd)
        for inverse in (False, True):
            out.zero_()
            launches = fft_b200.launch_count()
            (f.ifft if inverse else f.fft)(xd, out)
            torch.cuda.synchronize()
            assert fft_b200.launch_count() - launches == 1
            ref = oracle.run(oracle.KIND_C2C_INV if inverse else oracle.KIND_C2C_FWD, x, n, threads=8)[0]
            assert oracle.rel_l2(out.cpu().numpy(), ref) <= tol(n), (n, batch, inverse)
    f.fft(xd, xd)
    torch.cuda.synchronize()
    assert oracle.rel_l2(xd.cpu().numpy(), oracle.run(oracle.KIND_C2C_FWD, x, n, threads=8)[0]) <= tol(n)
