"""Multi-GPU tests (-m gpu; skipped on boxes with fewer than 2 GPUs): the distributed four-step with real ranks,
both exchange implementations (NCCL all_to_all_single, and the fused transpose + direct peer stores over NVLink),
checked against the oracle and against each other."""
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def test_distributed_four_step_real_ranks(oracle):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "gpu_multi_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    err = "\n".join(l for l in res.stderr.splitlines() if "Error" in l or "assert" in l or "File \"/root" in l)
    assert "MULTI-GPU-OK" in res.stdout, (res.stdout[-1500:] + "\n" + err[-2500:])


def test_single_process_distributed_plan_real_devices(oracle):
    """ssfft_dist_* (C ABI) over the GPUs of one process: peer stores over NVLink, chunked phases on two streams per
    device, the last exchange straight into the output shards.  Against the oracle at 2^16 ... 2^22, forward and inverse,
    natural and transposed output, repeated calls; needs at least 2 GPUs."""
    import numpy as np
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import math
    from fft_b200.dist import LocalDistFFT1D
    for world in sorted({2, min(4, torch.cuda.device_count()), min(8, torch.cuda.device_count())}):
        if world > torch.cuda.device_count():
            continue
        for n in (1 << 16, 1 << 22, 3 << 18):
            if n % (world * world):
                continue
            lim = 1e-6 * math.log2(n)
            plan = LocalDistFFT1D(n, list(range(world)))
            x = oracle.uniform_complex((1, n), 50 + world, np.complex64)
            per = n // world
            xs = [torch.from_numpy(x[0, r * per:(r + 1) * per].copy()).to(f"cuda:{r}") for r in range(world)]
            outs = [torch.zeros(per, dtype=torch.complex64, device=f"cuda:{r}") for r in range(world)]
            for _ in range(3):
                plan.fft(xs, outs)
            plan.synchronize()
            got = np.concatenate([o.cpu().numpy() for o in outs])[None, :]
            ref = oracle.run(oracle.KIND_C2C_FWD, x, n, threads=4)[0]
            assert oracle.rel_l2(got, ref) <= lim, (n, world, plan.describe())
            backs = [torch.zeros_like(o) for o in outs]
            plan.ifft(outs, backs)
            plan.synchronize()
            back = np.concatenate([o.cpu().numpy() for o in backs])[None, :]
            assert oracle.rel_l2(back / n, x) <= 2 * lim, (n, world)
            tp = LocalDistFFT1D(n, list(range(world)), transposed_output=True)
            touts = [torch.zeros(per, dtype=torch.complex64, device=f"cuda:{r}") for r in range(world)]
            tp.fft(xs, touts)
            tp.synchronize()
            t = np.concatenate([o.cpu().numpy() for o in touts]).reshape(tp.n1, tp.n2)
            assert oracle.rel_l2(t.T.reshape(1, n), ref) <= lim, (n, world, "transposed")
            plan.close()
            tp.close()
