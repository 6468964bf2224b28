# round 2, GPU call ab: fp64 single-pass kernels for 4608 / 6144 / 9216 -- parity and rate against the previous paths
set -x
mkdir -p gpurun_out
python - <<'P'
import numpy as np, torch, math, sys
sys.path.insert(0, '.')
import fft_b200
from oracle import oracle as O
for n in (4608, 6144, 9216):
    f = fft_b200.FFT(n, dtype="float64")
    for batch in (1, 3, 200):
        x = O.uniform_complex((batch, n), 5, np.complex128)
        xd = torch.from_numpy(x).cuda(); y = torch.empty_like(xd); z = torch.empty_like(xd)
        f.fft(xd, y); f.ifft(y, z); torch.cuda.synchronize()
        want = O.run(O.KIND_C2C_FWD, x, n, 4)[0]
        e1 = O.rel_l2(y.cpu().numpy(), want); e2 = O.rel_l2(z.cpu().numpy() / n, x)
        assert e1 <= 1e-14 * math.log2(n) and e2 <= 2e-14 * math.log2(n), (n, batch, e1, e2)
    r = fft_b200.RealFFT(2 * n, dtype="float64")
    xr = O.uniform(5 * 2 * n, 6, np.float64).reshape(5, 2 * n)
    xrd = torch.from_numpy(xr).cuda(); spec = torch.empty((5, n), dtype=torch.complex128, device="cuda"); back = torch.empty_like(xrd)
    r.fft(xrd, spec); r.ifft(spec, back); torch.cuda.synchronize()
    assert O.rel_l2(spec.cpu().numpy(), O.run(O.KIND_R2C, xr, 2 * n, 4)[0]) <= 1e-14 * math.log2(2 * n), (n, "r2c", r.describe())
    assert O.rel_l2(back.cpu().numpy() / (2 * n), xr) <= 2e-14 * math.log2(2 * n)
    print(n, "ok", f.describe()[:80])
print("F64-FUSED-OK")
P
timeout 300 python tools/sweep.py r02ab float64 4608 6144 9216 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02ab_f64.txt
SSFFT_DISABLE_FUSED=1 timeout 300 python tools/sweep.py r02ab_off float64 4608 6144 9216 2>&1 | grep "^N=" | sed "s/^/no fused  /" | tee -a gpurun_out/sweep_r02ab_f64.txt
