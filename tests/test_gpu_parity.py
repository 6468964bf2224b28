"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle.

Bar (BASELINE.json north_star): relative L2 error <= 1e-6 * log2(N) for float and 1e-14 * log2(N) for
double, per transform, against the reference algorithm in the SAME precision (oracle/ == reference
bit for bit, see test_oracle.py).  Sizes follow the reference's tests (tests/00-fft.cpp:8-16,
tests/01-real.cpp) plus every BASELINE config size; large batches are checked through
size-independent properties (fft->ifft == N x, linearity, Parseval) plus an oracle-checked subset.
"""
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import fft_b200  # noqa: E402

SEED = 7
REF_TEST_SIZES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 3, 6, 9, 12, 18, 24, 5, 10, 15, 20, 25, 7, 14, 21, 28, 49,
                  11, 13, 17, 19, 22, 23]
CONFIG_SIZES = [500, 512, 1000, 1024, 2048, 2187, 3125, 3000, 4096, 6000, 8192, 16384]
# the reference's fast sizes 2^k * {3, 9} (FFT::sizeMinimum/sizeMaximum): fused kernels with ragged passes
FAST_SIZES = [96, 192, 384, 768, 1536, 3072, 6144, 144, 288, 576, 1152, 2304, 4608, 9216]
LARGE_SIZES = [16384, 32768, 65536, 3 * 2 ** 15, 100000, 2 ** 17, 2 ** 18, 2 ** 19, 2 ** 20]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")


def tol(n, dtype):
    lg = max(1.0, math.log2(max(n, 2)))
    return (1e-6 if dtype in (np.complex64, np.float32) else 1e-14) * lg


def cdt_of(prec):
    return (np.complex64, torch.complex64) if prec == "float32" else (np.complex128, torch.complex128)


def rdt_of(prec):
    return (np.float32, torch.float32) if prec == "float32" else (np.float64, torch.float64)


def gpu_c2c(x_np, prec, inverse=False):
    n = x_np.shape[-1]
    fft = fft_b200.FFT(n, dtype=prec)
    x = torch.from_numpy(x_np).cuda()
    x_before = x.clone()
    out = torch.empty_like(x)
    (fft.ifft if inverse else fft.fft)(x, out)
    torch.cuda.synchronize()
    assert torch.equal(x, x_before), "input was changed"  # tests/00-fft.cpp:44-46
    return out.cpu().numpy(), fft


@pytest.mark.parametrize("prec", ["float32", "float64"])
def test_c2c_small_sizes_vs_oracle(oracle, cuda_device, prec):
    npdt, _ = cdt_of(prec)
    for n in REF_TEST_SIZES:
        for batch in (1, 3, 64):
            x = oracle.uniform_complex((batch, n), SEED, npdt)
            for inverse in (False, True):
                y, _ = gpu_c2c(x, prec, inverse)
                ref = oracle.ifft(x) if inverse else oracle.fft(x)
                err = oracle.rel_l2(y, ref)
                assert err <= tol(n, npdt), (n, batch, inverse, err)


@pytest.mark.parametrize("prec", ["float32", "float64"])
@pytest.mark.parametrize("n", CONFIG_SIZES + FAST_SIZES)
def test_c2c_config_sizes_vs_oracle(oracle, cuda_device, prec, n):
    npdt, _ = cdt_of(prec)
    batch = 37  # ragged against every transforms-per-block setting
    x = oracle.uniform_complex((batch, n), SEED, npdt)
    for inverse in (False, True):
        y, fft = gpu_c2c(x, prec, inverse)
        ref = oracle.run(oracle.KIND_C2C_INV if inverse else oracle.KIND_C2C_FWD, x, n, threads=8)[0]
        err = oracle.rel_l2(y, ref)
        assert err <= tol(n, npdt), (n, inverse, err, fft.describe())


@pytest.mark.parametrize("prec", ["float32", "float64"])
@pytest.mark.parametrize("n", LARGE_SIZES)
def test_c2c_large_four_step_vs_oracle(oracle, cuda_device, prec, n):
    npdt, _ = cdt_of(prec)
    x = oracle.uniform_complex((3, n), SEED, npdt)
    for inverse in (False, True):
        y, fft = gpu_c2c(x, prec, inverse)
        ref = oracle.run(oracle.KIND_C2C_INV if inverse else oracle.KIND_C2C_FWD, x, n, threads=3)[0]
        err = oracle.rel_l2(y, ref)
        assert err <= tol(n, npdt), (n, inverse, err, fft.describe())


def test_golden_vectors(oracle, cuda_device):
    """CUDA path vs outputs of the unmodified reference stored in tests/golden/."""
    g = np.load(GOLDEN)
    checked = 0
    for key in g.files:
        parts = key.split("_")
        prec = "float32" if parts[-2] == "f32" else "float64"
        n = int(parts[-1])
        if parts[0] == "c2c":
            npdt, _ = cdt_of(prec)
            x = oracle.uniform_complex((1, n), SEED, npdt)
            y, _ = gpu_c2c(x, prec, parts[1] == "inv")
            assert oracle.rel_l2(y, g[key][None]) <= tol(n, npdt), key
            checked += 1
        elif parts[0] in ("r2c", "m2c") and n in (2, 4, 6, 10, 30, 64, 98, 256, 1000, 4096):
            npdt, tdt = rdt_of(prec)
            x = oracle.uniform(n, SEED, npdt).reshape(1, n)
            cls = fft_b200.ModifiedRealFFT if parts[0] == "m2c" else fft_b200.RealFFT
            r = cls(n, dtype=prec)
            out = torch.empty((1, n // 2), dtype=cdt_of(prec)[1], device="cuda")
            r.fft(torch.from_numpy(x).cuda(), out)
            assert oracle.rel_l2(out.cpu().numpy(), g[key][None]) <= tol(n, npdt), key
            checked += 1
    assert checked > 100


@pytest.mark.parametrize("prec", ["float32", "float64"])
def test_individual_bins(cuda_device, prec):
    """tests/00-fft.cpp:19-60: input e^{+2 pi i n bin/N} -> N delta[bin], every bin at once as a batch."""
    npdt, _ = cdt_of(prec)
    for n in REF_TEST_SIZES + [1000, 4096]:
        k = np.arange(n)
        x = np.exp(2j * np.pi * ((np.outer(k, k)) % n) / n).astype(npdt)
        y, _ = gpu_c2c(x, prec)
        err = np.abs(y - n * np.eye(n)).max() / n
        assert err < (2e-6 if prec == "float32" else 1e-13) * max(1, math.log2(max(n, 2))), (n, err)


@pytest.mark.parametrize("prec", ["float32", "float64"])
def test_linearity_and_inverse_properties(oracle, cuda_device, prec):
    """tests/00-fft.cpp:62-145 on the GPU path."""
    npdt, _ = cdt_of(prec)
    for n in REF_TEST_SIZES + CONFIG_SIZES + [32768, 65536]:
        a = oracle.uniform_complex((2, n), 1, npdt)
        b = oracle.uniform_complex((2, n), 2, npdt)
        fa, _ = gpu_c2c(a, prec)
        fb, _ = gpu_c2c(b, prec)
        fab, _ = gpu_c2c((a + b).astype(npdt), prec)
        assert oracle.rel_l2(fab, fa + fb) <= 2 * tol(n, npdt), n
        back, _ = gpu_c2c(fa, prec, inverse=True)
        assert oracle.rel_l2(back, n * a) <= 2 * tol(n, npdt), n


@pytest.mark.parametrize("prec", ["float32", "float64"])
@pytest.mark.parametrize("cls_name", ["RealFFT", "ModifiedRealFFT"])
def test_real_vs_oracle(oracle, cuda_device, prec, cls_name):
    """tests/01-real.cpp: even sizes 2..98 (+ config sizes), packing of bin 0, inverse scaled by N."""
    npdt, tdt = rdt_of(prec)
    _, tcdt = cdt_of(prec)
    modified = cls_name == "ModifiedRealFFT"
    cls = getattr(fft_b200, cls_name)
    for n in list(range(2, 100, 2)) + [128, 256, 1000, 2048, 4096, 8192, 6000, 192, 1536, 3072, 4608, 32768, 65536, 2 ** 17, 2 ** 18, 2 ** 20]:
        batch = 5
        x = oracle.uniform(batch * n, SEED, npdt).reshape(batch, n)
        r = cls(n, dtype=prec)
        assert r.size() == n
        xd = torch.from_numpy(x).cuda()
        # output buffer is a full N complex per transform filled with a sentinel: [N/2, N) must stay untouched
        full = torch.full((batch * (n // 2) + n,), 123.0 + 456.0j, dtype=tcdt, device="cuda")
        out = full[: batch * (n // 2)].view(batch, n // 2)
        r.fft(xd, out)
        torch.cuda.synchronize()
        assert torch.all(full[batch * (n // 2):] == 123.0 + 456.0j), "wrote past the N/2 bins"  # tests/01-real.cpp:70-73
        assert torch.equal(xd, torch.from_numpy(x).cuda()), "input was changed"
        ref = oracle.rfft(x, modified)
        err = oracle.rel_l2(out.cpu().numpy(), ref)
        assert err <= tol(n, npdt), (cls_name, n, err, r.describe())
        back = torch.empty((batch, n), dtype=tdt, device="cuda")
        r.ifft(out, back)
        ref_back = oracle.irfft(ref, modified)
        err = oracle.rel_l2(back.cpu().numpy(), ref_back)
        assert err <= 2 * tol(n, npdt), (cls_name, n, "inverse", err)
        assert oracle.rel_l2(back.cpu().numpy(), n * x) <= 2 * tol(n, npdt)


def test_generic_path_matches_too(oracle, cuda_device):
    """Force every size through the generic interpreter kernel (fresh process so the registry is empty)."""
    import subprocess
    import sys
    code = r"""
import os, sys, math
os.environ["SSFFT_DISABLE_FUSED"] = "1"
os.environ["SSFFT_DISABLE_TILED"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np, torch, fft_b200
from oracle import oracle as O
for prec, npdt in (("float32", np.complex64), ("float64", np.complex128)):
    for n in [64, 256, 1000, 1024, 2187, 3125, 4096, 6000, 65536]:
        x = O.uniform_complex((9 if n < 60000 else 2, n), 7, npdt)
        f = fft_b200.FFT(n, dtype=prec)
        assert "generic" in f.describe() and "fused" not in f.describe(), f.describe()
        xd = torch.from_numpy(x).cuda(); out = torch.empty_like(xd)
        f.fft(xd, out)
        err = O.rel_l2(out.cpu().numpy(), O.fft(x))
        lim = (1e-6 if prec == "float32" else 1e-14) * math.log2(n)
        assert err <= lim, (prec, n, err)
print("GENERIC-OK")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert "GENERIC-OK" in res.stdout, res.stdout + res.stderr


def test_api_surface(oracle, cuda_device):
    """setSize / setSizeMinimum / setSizeMaximum / size / fastDirection, and RealFFT's quirks."""
    f = fft_b200.FFT(1000, 1)
    assert f.size() == 1024
    assert fft_b200.FFT(1000, -1).size() == 768
    assert f.setSize(96) == 96 and f.size() == 96
    assert f.setSizeMinimum(1025) == 1152 and f.setSizeMaximum(1025) == 1024
    r = fft_b200.RealFFT(64)
    assert r.setSize(256) == 128 and r.size() == 256
    assert fft_b200.RealFFT(7).size() == 6
    assert fft_b200.RealFFT(1000, 1).size() == 1026
    # empty batch and size-1 transform
    x = torch.zeros((0, 96), dtype=torch.complex64, device="cuda")
    f.setSize(96)
    f.fft(x, torch.empty_like(x))
    one = fft_b200.FFT(1)
    a = torch.tensor([[1.5 - 2j], [0.25 + 1j]], dtype=torch.complex64, device="cuda")
    b = torch.empty_like(a)
    one.fft(a, b)
    assert torch.equal(a, b)


def test_host_pointer_path(oracle, cuda_device):
    """The call a reference user makes: fft(host_in, host_out) -- staged H2D/D2H inside the library."""
    for prec, npdt in (("float32", np.complex64), ("float64", np.complex128)):
        for n, batch in ((24, 1), (4096, 64), (1000, 10)):
            x = oracle.uniform_complex((batch, n), 9, npdt)
            x0 = x.copy()
            out = np.empty_like(x)
            f = fft_b200.FFT(n, dtype=prec)
            f.fft(x, out)
            assert np.array_equal(x, x0)
            assert oracle.rel_l2(out, oracle.fft(x)) <= tol(n, npdt)
            back = np.empty_like(x)
            f.ifft(out, back)
            assert oracle.rel_l2(back, n * x) <= 2 * tol(n, npdt)
    xr = oracle.uniform(6 * 512, 3, np.float32).reshape(6, 512)
    r = fft_b200.RealFFT(512)
    spec = np.empty((6, 256), np.complex64)
    r.fft(xr, spec)
    assert oracle.rel_l2(spec, oracle.rfft(xr)) <= tol(512, np.float32)


def test_headline_batch_properties(oracle, cuda_device):
    """BASELINE config 2 at full size: N=4096 x 65536 fp32.  Oracle-checked subset + fft->ifft == N x and
    Parseval over the whole batch (size-independent properties)."""
    n, batch = 4096, 65536
    x = torch.empty((batch, n), dtype=torch.complex64, device="cuda")
    fft_b200.fill_uniform(x, 20261017)
    f = fft_b200.FFT(n)
    y = torch.empty_like(x)
    f.fft(x, y)
    idx = [0, 1, 2, 777, 32767, 32768, 65534, 65535]
    xs = x[idx].cpu().numpy()
    # device generator == host generator (so the oracle sees the same inputs without a transfer)
    for i, b in enumerate(idx):
        assert np.array_equal(xs[i], oracle.uniform_complex((n,), 20261017, np.complex64, first_idx=2 * n * b))
    assert oracle.rel_l2(y[idx].cpu().numpy(), oracle.fft(xs)) <= tol(n, np.complex64)
    # Parseval: sum |X|^2 == N sum |x|^2, per transform
    ex = (x.real.double() ** 2 + x.imag.double() ** 2).sum(dim=1)
    ey = (y.real.double() ** 2 + y.imag.double() ** 2).sum(dim=1)
    assert torch.max(torch.abs(ey / (n * ex) - 1)).item() < 1e-5
    z = torch.empty_like(x)
    f.ifft(y, z)
    num = torch.linalg.vector_norm((z - n * x).reshape(batch, -1), dim=1)
    den = torch.linalg.vector_norm((n * x).reshape(batch, -1), dim=1)
    assert (num / den).max().item() <= 2 * tol(n, np.complex64)


def test_distributed_four_step_logical_ranks_on_one_gpu(oracle, cuda_device):
    """The multi-GPU four-step code path with P logical ranks on one device (block swaps instead of NCCL):
    same kernels, same index logic as the real distributed run (fft_b200/dist.py)."""
    from fft_b200.dist import DistFFT1D

    for n, world in ((1 << 16, 2), (1 << 20, 4), (1 << 22, 8), (3 * (1 << 16), 2)):
        x = oracle.uniform_complex((n,), 11, np.complex64)
        plan = DistFFT1D(n, world, dtype=torch.complex64)
        per = n // world
        xs = [torch.from_numpy(x[r * per:(r + 1) * per].copy()).cuda() for r in range(world)]
        ys = plan.run_logical(xs)
        y = torch.cat(ys).cpu().numpy()
        ref = oracle.run(oracle.KIND_C2C_FWD, x[None], n, threads=1)[0]
        assert oracle.rel_l2(y[None], ref) <= tol(n, np.complex64), (n, world, plan.n1, plan.n2)
        back = torch.cat(plan.run_logical(ys, inverse=True)).cpu().numpy()
        assert oracle.rel_l2(back[None], (n * x)[None]) <= 2 * tol(n, np.complex64)
        assert plan.exchanges == 6


def test_exchange_building_blocks(oracle, cuda_device):
    """ssfft_transpose_twiddle / ssfft_permute102 (local steps of the distributed four-step) against numpy."""
    import ctypes

    from fft_b200 import _lib as L
    lib = L.load()
    for dt, prec, rtol in ((np.complex64, L.SSFFT_F32, 3e-7), (np.complex128, L.SSFFT_F64, 1e-15)):
        for rows, cols, row0, n_total, inv in ((48, 80, 0, 0, 0), (300, 70, 5, 300 * 70 * 3, 0), (257, 33, 1000, 1 << 30, 1)):
            x = oracle.uniform_complex((rows, cols), 4, dt)
            xd = torch.from_numpy(x).cuda()
            yd = torch.empty((cols, rows), dtype=xd.dtype, device="cuda")
            L.check(lib.ssfft_transpose_twiddle(xd.data_ptr(), yd.data_ptr(), 1, rows, cols, row0, n_total, inv, prec, None),
                    "ssfft_transpose_twiddle")
            ref = x.astype(np.complex128)
            if n_total:
                q = (np.arange(row0, row0 + rows, dtype=object)[:, None] * np.arange(cols, dtype=object)[None, :]) % n_total
                ref = ref * np.exp((2j if inv else -2j) * np.pi * q.astype(np.float64) / n_total)
            assert oracle.rel_l2(yd.cpu().numpy().reshape(1, -1), ref.T.reshape(1, -1)) <= rtol, (dt, rows, cols, n_total)
        a, b, run = 3, 5, 14
        x = oracle.uniform_complex((a, b, run), 6, dt)
        xd = torch.from_numpy(x).cuda()
        yd = torch.empty((b, a, run), dtype=xd.dtype, device="cuda")
        L.check(lib.ssfft_permute102(xd.data_ptr(), yd.data_ptr(), a, b, run, prec, None), "ssfft_permute102")
        assert np.array_equal(yd.cpu().numpy(), x.transpose(1, 0, 2))


def test_cluster_resident_four_step(oracle, cuda_device):
    """cluster.cuh: a transform lives in the shared memory of a thread-block cluster (DSMEM all-to-all, st.async +
    mbarrier, TMA tensor prefetch).  tests/gpu_cluster_check.py runs batches several times larger than the number of
    co-resident clusters (every cluster loops over many transforms, exercising the barrier protocol), C2C both
    directions and RealFFT forward / inverse, for every registered geometry (SSFFT_DSMEM_ALL=1), with and without TMA."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for tma in ("1", "0"):
        env = dict(os.environ, SSFFT_DSMEM_ALL="1", SSFFT_CLUSTER_TMA=tma, SSFFT_DISABLE_FLAT="1")
        res = subprocess.run([sys.executable, os.path.join(root, "tests", "gpu_cluster_check.py")], cwd=root, env=env,
                             capture_output=True, text=True, timeout=900)
        assert "CLUSTER-GPU-OK" in res.stdout, (tma, res.stdout[-2000:] + res.stderr[-2000:])
    # default planner choice since round 2: the ticket-queue four-step (flat.cuh); the cluster-resident kernel is the
    # fallback for inputs a tensor map cannot describe
    f = fft_b200.FFT(32768)
    assert "ticket-queue" in f.describe(), f.describe()


def test_l2_scratch_four_step_still_matches(oracle, cuda_device):
    """With the cluster-resident kernels disabled the same sizes run through the L2-scratch four-step (tiled.cuh)."""
    import subprocess
    import sys
    code = r"""
import os, sys, math
os.environ["SSFFT_DISABLE_DSMEM"] = "1"
os.environ["SSFFT_DISABLE_FLAT"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np, torch, fft_b200
from oracle import oracle as O
for n in (32768, 65536):
    x = O.uniform_complex((5, n), 7, np.complex64)
    f = fft_b200.FFT(n)
    assert "cluster-resident" not in f.describe(), f.describe()
    xd = torch.from_numpy(x).cuda(); out = torch.empty_like(xd)
    f.fft(xd, out)
    assert O.rel_l2(out.cpu().numpy(), O.fft(x)) <= 1e-6 * math.log2(n), n
    xr = O.uniform(3 * 2 * n, 8, np.float32).reshape(3, 2 * n)
    r = fft_b200.RealFFT(2 * n)
    assert "cluster-resident" not in r.describe(), r.describe()
    sp = torch.empty((3, n), dtype=torch.complex64, device="cuda")
    r.fft(torch.from_numpy(xr).cuda(), sp)
    assert O.rel_l2(sp.cpu().numpy(), O.rfft(xr)) <= 1e-6 * math.log2(2 * n), n
print("L2-FOURSTEP-OK")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert "L2-FOURSTEP-OK" in res.stdout, res.stdout + res.stderr


def test_pointers_offset_by_one_element(oracle, cuda_device):
    """Input and output that are only 8-byte aligned (a view starting at element 1): the TMA paths (bulk prefetch of the
    fused kernels, tensor prefetch of the cluster-resident kernel) need 16-byte aligned sources and must fall back to
    plain loads instead of faulting."""
    for n in (1024, 4096, 8192, 32768, 65536):
        batch = 9
        x = oracle.uniform_complex((batch, n), SEED, np.complex64)
        buf_in = torch.zeros(batch * n + 1, dtype=torch.complex64, device="cuda")
        buf_out = torch.zeros(batch * n + 1, dtype=torch.complex64, device="cuda")
        xin = buf_in[1:].view(batch, n)
        xin.copy_(torch.from_numpy(x))
        out = buf_out[1:].view(batch, n)
        assert xin.data_ptr() % 16 == 8 and out.data_ptr() % 16 == 8
        f = fft_b200.FFT(n)
        f.fft(xin, out)
        torch.cuda.synchronize()
        assert oracle.rel_l2(out.cpu().numpy(), oracle.run(oracle.KIND_C2C_FWD, x, n, threads=4)[0]) <= tol(n, np.complex64), n
    for n in (2048, 8192, 65536):
        batch = 5
        xr = oracle.uniform(batch * n, SEED, np.float32).reshape(batch, n)
        buf_in = torch.zeros(batch * n + 2, dtype=torch.float32, device="cuda")
        xin = buf_in[2:].view(batch, n)
        xin.copy_(torch.from_numpy(xr))
        assert xin.data_ptr() % 16 == 8
        spec = torch.empty((batch, n // 2), dtype=torch.complex64, device="cuda")
        r = fft_b200.RealFFT(n)
        r.fft(xin, spec)
        torch.cuda.synchronize()
        assert oracle.rel_l2(spec.cpu().numpy(), oracle.rfft(xr)) <= tol(n, np.float32), n


def test_in_place_device_transforms(oracle, cuda_device):
    """The reference is out-of-place only (input must not alias output); the device path also accepts in == out:
    every kernel has consumed a transform's input before it stores the first output of that transform."""
    for n in (64, 1000, 4096, 16384, 32768, 65536, 2 ** 18):
        batch = 7
        x = oracle.uniform_complex((batch, n), SEED, np.complex64)
        xd = torch.from_numpy(x).cuda()
        f = fft_b200.FFT(n)
        f.fft(xd, xd)
        torch.cuda.synchronize()
        ref = oracle.run(oracle.KIND_C2C_FWD, x, n, threads=4)[0]
        assert oracle.rel_l2(xd.cpu().numpy(), ref) <= tol(n, np.complex64), (n, f.describe())
    for n in (256, 4096, 65536):
        batch = 5
        xr = oracle.uniform(batch * n, SEED, np.float32).reshape(batch, n)
        buf = torch.from_numpy(xr).cuda()
        r = fft_b200.RealFFT(n)
        spec = buf.view(torch.complex64)  # same bytes: N reals -> N/2 complex per transform
        r.fft(buf, spec)
        torch.cuda.synchronize()
        assert oracle.rel_l2(spec.cpu().numpy(), oracle.rfft(xr)) <= tol(n, np.float32), (n, r.describe())
