# round 2, GPU call aa: single-pass kernels for 32 / 36 / 48 / 72 -- parity and rate
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 300 python tools/sweep.py r02aa float32 32 36 48 72 64 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02aa_f32.txt
timeout 300 python tools/sweep.py r02aa float64 32 36 48 72 64 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02aa_f64.txt
python - <<'P'
import numpy as np, torch, math, sys
sys.path.insert(0, '.')
import fft_b200
from oracle import oracle as O
for prec, npdt, eps in (("float32", np.complex64, 1e-6), ("float64", np.complex128, 1e-14)):
    for n in (32, 36, 48, 72):
        f = fft_b200.FFT(n, dtype=prec)
        for batch in (1, 31, 1000):
            x = O.uniform_complex((batch, n), 5, npdt)
            xd = torch.from_numpy(x).cuda(); y = torch.empty_like(xd); z = torch.empty_like(xd)
            f.fft(xd, y); f.ifft(y, z); torch.cuda.synchronize()
            want = O.run(O.KIND_C2C_FWD, x, n, 2)[0]
            e1 = O.rel_l2(y.cpu().numpy(), want); e2 = O.rel_l2(z.cpu().numpy() / n, x)
            assert e1 <= eps * math.log2(n) and e2 <= 2 * eps * math.log2(n), (prec, n, batch, e1, e2)
        r = fft_b200.RealFFT(2 * n, dtype=prec)
        xr = O.uniform(7 * 2 * n, 6, np.float32 if prec == "float32" else np.float64).reshape(7, 2 * n)
        xrd = torch.from_numpy(xr).cuda(); spec = torch.empty((7, n), dtype=xd.dtype, device="cuda"); back = torch.empty_like(xrd)
        r.fft(xrd, spec); r.ifft(spec, back); torch.cuda.synchronize()
        assert O.rel_l2(spec.cpu().numpy(), O.run(O.KIND_R2C, xr, 2 * n, 2)[0]) <= eps * math.log2(2 * n), (prec, n, "r2c", r.describe())
        assert O.rel_l2(back.cpu().numpy() / (2 * n), xr) <= 2 * eps * math.log2(2 * n)
        print(prec, n, "ok", f.describe()[:70], "|", r.describe()[:60])
print("SMALL-FUSED-OK")
P
