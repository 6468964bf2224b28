"""GPU tests (-m gpu) of what round 2 added around the kernels: lengths the reference accepts and round 1 rejected
(Bluestein), the rebuilt host-pointer path (ring of slices on three plan-owned streams, zero-copy staging for tiny calls)
and the ordering of calls that share one plan.  Everything goes through the C ABI (fft_b200 is a ctypes mirror)."""
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import fft_b200  # noqa: E402


def tol(n, dtype):
    return (1e-6 if dtype in (np.complex64, np.float32) else 1e-14) * math.log2(n)


def exact_dft(x):
    """numpy's double-precision FFT of the (exactly representable) input: the 'exact transform' at these tolerances."""
    return np.fft.fft(x.astype(np.complex128), axis=-1)


@pytest.mark.parametrize("prec", ["float32", "float64"])
@pytest.mark.parametrize("n", [16411, 2 * 65537, 3 * 16411])
def test_large_prime_factor_lengths(oracle, cuda_device, prec, n):
    """signalsmith-fft.h:146-150, 187-215: the reference runs a remaining large prime as one O(p^2) step, so every length
    works there.  Here such lengths go through Bluestein; bar = the parity bar against the exact transform."""
    npdt = np.complex64 if prec == "float32" else np.complex128
    f = fft_b200.FFT(n, dtype=prec)
    assert "Bluestein" in f.describe(), f.describe()
    x = oracle.uniform_complex((3, n), 21, npdt)
    xd = torch.from_numpy(x).cuda()
    y = torch.empty_like(xd)
    z = torch.empty_like(xd)
    f.fft(xd, y)
    f.ifft(y, z)
    torch.cuda.synchronize()
    want = exact_dft(x)
    err = oracle.rel_l2(y.cpu().numpy().astype(np.complex128), want)
    assert err <= tol(n, npdt), (n, prec, err)
    rt = oracle.rel_l2(z.cpu().numpy().astype(np.complex128) / n, x.astype(np.complex128))
    assert rt <= 2 * tol(n, npdt), (n, prec, rt)
    if n == 16411:  # no worse than the reference itself (its float phases are rounded before cos / sin)
        ref = oracle.run(oracle.KIND_C2C_FWD, x[:1], n, threads=1)[0]
        ref_err = oracle.rel_l2(ref.astype(np.complex128), want[:1])
        assert err <= max(tol(n, npdt), 1.5 * ref_err), (err, ref_err)


@pytest.mark.parametrize("prec", ["float32", "float64"])
def test_real_transform_with_large_prime_factor(oracle, cuda_device, prec):
    n = 2 * 16411
    rdt = np.float32 if prec == "float32" else np.float64
    cdt = np.complex64 if prec == "float32" else np.complex128
    r = fft_b200.RealFFT(n, dtype=prec)
    assert "Bluestein" in r.describe(), r.describe()
    x = oracle.uniform(2 * n, 22, rdt).reshape(2, n)
    xd = torch.from_numpy(x).cuda()
    spec = torch.empty((2, n // 2), dtype=torch.complex64 if prec == "float32" else torch.complex128, device="cuda")
    back = torch.empty_like(xd)
    r.fft(xd, spec)
    r.ifft(spec, back)
    torch.cuda.synchronize()
    full = np.fft.fft(x.astype(np.float64), axis=-1)
    want = full[:, : n // 2].copy()
    want[:, 0] = full[:, 0].real + 1j * full[:, n // 2].real  # (DC, Nyquist) packed in bin 0 (:459-462)
    err = oracle.rel_l2(spec.cpu().numpy().astype(np.complex128), want)
    assert err <= tol(n, cdt), err
    assert oracle.rel_l2(back.cpu().numpy().astype(np.float64) / n, x.astype(np.float64)) <= 2 * tol(n, cdt)


def test_host_path_slices_match_device_path(oracle, cuda_device):
    """ssfft_exec_host: a call large enough to be cut into slices (ring of three buffers, three streams) gives the same
    bits as the device path, from pinned and from pageable host memory; so does the zero-copy path of tiny calls."""
    for n, batch, prec in ((4096, 16384, "float32"), (65536, 600, "float32"), (1000, 9000, "float64")):
        npdt = np.complex64 if prec == "float32" else np.complex128
        f = fft_b200.FFT(n, dtype=prec)
        xd = torch.empty((batch, n), dtype=torch.complex64 if prec == "float32" else torch.complex128, device="cuda")
        fft_b200.fill_uniform(xd, 5)
        yd = torch.empty_like(xd)
        f.fft(xd, yd)
        torch.cuda.synchronize()
        want = yd.cpu().numpy()
        x_pageable = xd.cpu().numpy()
        out = np.empty_like(x_pageable)
        f.fft(x_pageable, out)
        assert np.array_equal(out, want), (n, "pageable")
        x_pinned = torch.empty(xd.shape, dtype=xd.dtype, pin_memory=True)
        x_pinned.copy_(xd)
        out_pinned = torch.empty(xd.shape, dtype=xd.dtype, pin_memory=True)
        f.fft(x_pinned.numpy(), out_pinned.numpy())
        assert np.array_equal(out_pinned.numpy(), want), (n, "pinned")
        f.ifft(x_pageable, out)
        f.ifft(xd, yd)
        torch.cuda.synchronize()
        assert np.array_equal(out, yd.cpu().numpy()), (n, "inverse")
        del xd, yd
    # tiny calls: kernels read and write pinned mapped staging directly
    for n, batch, prec in ((1024, 1, "float64"), (4096, 4, "float32"), (256, 3, "float32"), (16384, 1, "float32")):
        npdt = np.complex64 if prec == "float32" else np.complex128
        f = fft_b200.FFT(n, dtype=prec)
        x = oracle.uniform_complex((batch, n), 9, npdt)
        out = np.empty_like(x)
        for _ in range(3):
            f.fft(x, out)
        yd = torch.empty((batch, n), dtype=torch.from_numpy(x).dtype, device="cuda")
        f.fft(torch.from_numpy(x).cuda(), yd)
        torch.cuda.synchronize()
        assert np.array_equal(out, yd.cpu().numpy()), (n, batch, "zero copy")
        assert oracle.rel_l2(out, oracle.run(oracle.KIND_C2C_FWD, x, n, threads=1)[0]) <= tol(n, npdt)
    # real transforms through the host path
    n, batch = 4096, 40000
    r = fft_b200.RealFFT(n)
    xr = oracle.uniform(batch * n, 4, np.float32).reshape(batch, n)
    spec = np.empty((batch, n // 2), dtype=np.complex64)
    r.fft(xr, spec)
    sd = torch.empty((batch, n // 2), dtype=torch.complex64, device="cuda")
    r.fft(torch.from_numpy(xr).cuda(), sd)
    torch.cuda.synchronize()
    assert np.array_equal(spec, sd.cpu().numpy())
    back = np.empty_like(xr)
    r.ifft(spec, back)
    assert oracle.rel_l2(back / n, xr) <= 2 * tol(n, np.float32)


def test_calls_on_one_plan_from_two_streams_are_ordered(oracle, cuda_device):
    """A plan owns scratch (four-step sizes): calls issued on different streams must not overlap on the device."""
    n, batch = 65536, 256
    f = fft_b200.FFT(n)
    xs = [torch.empty((batch, n), dtype=torch.complex64, device="cuda") for _ in range(2)]
    for i, x in enumerate(xs):
        fft_b200.fill_uniform(x, 30 + i)
    want = []
    for x in xs:
        y = torch.empty_like(x)
        f.fft(x, y)
        torch.cuda.synchronize()
        want.append(y)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [torch.empty_like(x) for x in xs]
    torch.cuda.synchronize()
    for rep in range(4):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                f.fft(xs[i], outs[i])
    torch.cuda.synchronize()
    for i in range(2):
        assert torch.equal(outs[i], want[i]), i


@pytest.mark.parametrize("world", [1, 2, 4])
def test_single_process_distributed_plan_logical_ranks(oracle, cuda_device, world):
    """ssfft_dist_plan_create / ssfft_dist_exec_c2c (C ABI of BASELINE config 5) with every logical rank on GPU 0: the same
    kernels, events and chunked phases as on several GPUs, checked against the oracle; repeated calls reuse the buffers."""
    from fft_b200.dist import LocalDistFFT1D

    for n in (1 << 14, 1 << 18, 3 << 16):
        if n % (world * world):
            continue
        plan = LocalDistFFT1D(n, [0] * world)
        x = oracle.uniform_complex((1, n), 40 + world, np.complex64)
        per = n // world
        xs = [torch.from_numpy(x[0, r * per:(r + 1) * per].copy()).cuda() for r in range(world)]
        outs = [torch.zeros(per, dtype=torch.complex64, device="cuda") for _ in range(world)]
        for rep in range(2):
            plan.fft(xs, outs)
        plan.synchronize()
        got = np.concatenate([o.cpu().numpy() for o in outs])[None, :]
        ref = oracle.run(oracle.KIND_C2C_FWD, x, n, threads=1)[0]
        assert oracle.rel_l2(got, ref) <= tol(n, np.complex64), (n, world, plan.describe())
        backs = [torch.zeros_like(o) for o in outs]
        plan.ifft(outs, backs)
        plan.synchronize()
        back = np.concatenate([o.cpu().numpy() for o in backs])[None, :]
        assert oracle.rel_l2(back / n, x) <= 2 * tol(n, np.complex64), (n, world)
        # transposed output: shard r = rows k1 in [r a, (r+1) a) of X[k1 + N1 k2], laid out [a][N2]
        tp = LocalDistFFT1D(n, [0] * world, transposed_output=True)
        touts = [torch.zeros(per, dtype=torch.complex64, device="cuda") for _ in range(world)]
        tp.fft(xs, touts)
        tp.synchronize()
        t = np.concatenate([o.cpu().numpy() for o in touts]).reshape(tp.n1, tp.n2)  # [k1][k2]
        assert oracle.rel_l2(t.T.reshape(1, n), ref) <= tol(n, np.complex64), (n, world, "transposed")
        plan.close()
        tp.close()


@pytest.mark.parametrize("prec", ["float32", "float64"])
@pytest.mark.parametrize("n", [9 << 19, 1 << 23, 3 << 21, 1 << 24])
def test_composite_lengths_vs_oracle(oracle, cuda_device, prec, n):
    """The reference's benchmark set (benchmark/benchmark.h:27-52: 2^k x {1, 3, 9} up to 2^24) beyond the specialised
    kernels: one radix pass around a fast plan (composite.cuh).  Against the oracle, forward and inverse, batch > 1 for the
    shorter lengths, plus the real transform of twice the length."""
    if prec == "float64" and n > (1 << 22):
        pytest.skip("covered in float32; keeps the suite short")
    npdt = np.complex64 if prec == "float32" else np.complex128
    f = fft_b200.FFT(n, dtype=prec)
    assert "composite" in f.describe(), f.describe()
    batch = 3 if n <= (1 << 20) else 1
    x = oracle.uniform_complex((batch, n), 31, npdt)
    xd = torch.from_numpy(x).cuda()
    y = torch.empty_like(xd)
    z = torch.empty_like(xd)
    f.fft(xd, y)
    f.ifft(y, z)
    torch.cuda.synchronize()
    want = oracle.run(oracle.KIND_C2C_FWD, x, n, threads=8)[0]
    err = oracle.rel_l2(y.cpu().numpy(), want)
    assert err <= tol(n, npdt), (n, prec, err, f.describe())
    want_inv = oracle.run(oracle.KIND_C2C_INV, want, n, threads=8)[0]
    assert oracle.rel_l2(z.cpu().numpy(), want_inv) <= 2 * tol(n, npdt), (n, prec)
    assert torch.equal(xd.cpu(), torch.from_numpy(x)), "input was modified"
    if n <= (1 << 21):
        rdt = np.float32 if prec == "float32" else np.float64
        r = fft_b200.RealFFT(2 * n, dtype=prec)
        xr = oracle.uniform(2 * n, 32, rdt).reshape(1, 2 * n)
        xrd = torch.from_numpy(xr).cuda()
        spec = torch.empty((1, n), dtype=xd.dtype, device="cuda")
        back = torch.empty_like(xrd)
        r.fft(xrd, spec)
        r.ifft(spec, back)
        torch.cuda.synchronize()
        want_r = oracle.run(oracle.KIND_R2C, xr, 2 * n, threads=8)[0]
        assert oracle.rel_l2(spec.cpu().numpy(), want_r) <= tol(2 * n, npdt), (2 * n, prec, r.describe())
        assert oracle.rel_l2(back.cpu().numpy() / (2 * n), xr) <= 2 * tol(2 * n, npdt)


@pytest.mark.parametrize("prec", ["float32", "float64"])
def test_tiny_lengths_one_thread_per_transform(oracle, cuda_device, prec):
    """N <= 24 (tiny.cuh): every supported length, forward / inverse / in place, batches that are not multiples of a CTA's
    256 transforms, against the oracle; the first sizes of the reference's benchmark (benchmark/benchmark.h:27-52)."""
    npdt = np.complex64 if prec == "float32" else np.complex128
    eps = 1e-6 if prec == "float32" else 1e-14
    for n in list(range(1, 17)) + [18, 20, 24]:
        f = fft_b200.FFT(n, dtype=prec)
        assert "one thread per transform" in f.describe(), f.describe()
        for batch in (1, 255, 257, 5000):
            x = oracle.uniform_complex((batch, n), 60 + n, npdt)
            xd = torch.from_numpy(x).cuda()
            y = torch.empty_like(xd)
            z = torch.empty_like(xd)
            launches = fft_b200.launch_count()
            f.fft(xd, y)
            assert fft_b200.launch_count() - launches == 1
            f.ifft(y, z)
            torch.cuda.synchronize()
            lim = eps * max(1.0, math.log2(n))
            want = oracle.run(oracle.KIND_C2C_FWD, x, n, threads=2)[0]
            assert oracle.rel_l2(y.cpu().numpy(), want) <= lim, (n, batch, prec)
            assert oracle.rel_l2(z.cpu().numpy(), oracle.run(oracle.KIND_C2C_INV, want, n, threads=2)[0]) <= 2 * lim, (n, batch, prec)
            assert torch.equal(xd.cpu(), torch.from_numpy(x))
        xi = torch.from_numpy(x).cuda()
        f.fft(xi, xi)
        torch.cuda.synchronize()
        assert oracle.rel_l2(xi.cpu().numpy(), want) <= lim, (n, "in place")
