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
