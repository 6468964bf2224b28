"""tools/c2r_check.py -- RealFFT inverse of 65536 on the ticket-queue kernels, small batch, against the forward result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fft_b200
if os.environ.get("SSFFT_LIB"):
    from fft_b200 import _lib as _L
    _L.LIB_PATH = fft_b200.LIB_PATH = os.environ["SSFFT_LIB"]
n, batch = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 3
r = fft_b200.RealFFT(n)
print(r.describe())
x = torch.empty((batch, n), dtype=torch.float32, device="cuda")
fft_b200.fill_uniform(x, 1)
spec = torch.empty((batch, n // 2), dtype=torch.complex64, device="cuda")
back = torch.empty_like(x)
r.fft(x, spec)
torch.cuda.synchronize()
print("forward ok")
r.ifft(spec, back)
torch.cuda.synchronize()
print("inverse ok, round trip error", float(torch.linalg.vector_norm(back / n - x) / torch.linalg.vector_norm(x)))
