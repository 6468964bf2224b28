"""CPU tests (no GPU) of the multi-GPU host logic: world_size-2 gloo processes run fft_b200.dist.DistFFT1D with a
numpy stand-in for the local GPU steps (the rank / index / exchange logic is what is under test), and the batch
sharding helper.  Results are checked against the oracle on the gathered vector."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fft_b200.dist import DistFFT1D, choose_factors, shard_range


class NumpyOps:
    """Test-only stand-in for fft_b200.dist.CudaOps (same interface, CPU tensors)."""

    def fft_rows(self, x, inverse, out):
        a = x.numpy()
        y = np.fft.ifft(a, axis=-1) * a.shape[-1] if inverse else np.fft.fft(a, axis=-1)
        out.view(x.shape).copy_(torch.from_numpy(y.astype(a.dtype)))
        return out.view(x.shape)

    def transpose(self, x, out, row0=0, n_total=0, inverse=False):
        rows, cols = x.shape
        a = x.numpy().astype(np.complex128)
        if n_total:
            q = (np.arange(row0, row0 + rows, dtype=np.int64)[:, None] * np.arange(cols, dtype=np.int64)[None, :]) % n_total
            a = a * np.exp((2j if inverse else -2j) * np.pi * q / n_total)
        out.view(cols, rows).copy_(torch.from_numpy(np.ascontiguousarray(a.T).astype(x.numpy().dtype)))
        return out.view(cols, rows)

    def permute102(self, x, out):
        a, b, run = x.shape
        out.view(b, a, run).copy_(x.permute(1, 0, 2))
        return out.view(b, a, run)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, n1, inverse, transposed, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        x = (rng.random(n) - 0.5 + 1j * (rng.random(n) - 0.5)).astype(np.complex128)
        per = n // world
        plan = DistFFT1D(n, world, rank, dtype=torch.complex128, ops=NumpyOps(), n1=n1, transposed_output=transposed)
        xl = torch.from_numpy(x[rank * per:(rank + 1) * per].copy())
        out = torch.empty_like(xl)
        (plan.ifft if inverse else plan.fft)(xl, out)
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        if rank == 0:
            ret["y"] = torch.cat(gathered).numpy()
            ret["x"] = x
            ret["exchanges"] = plan.exchanges
            ret["shape"] = (plan.n1, plan.n2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,n1,inverse,transposed", [(64, 8, False, False), (96, 8, True, False), (1024, None, False, False),
                                                     (240, 12, False, True)])
def test_distributed_four_step_gloo_world2(oracle, n, n1, inverse, transposed):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, n1, inverse, transposed, ret), nprocs=world, join=True)
    x, y = ret["x"], ret["y"]
    ref = oracle.ifft(x[None])[0] if inverse else oracle.fft(x[None])[0]
    if transposed:
        n1_, n2_ = ret["shape"]
        # rank r holds [k1 local][k2] = X[k1 + n1*k2]; gathered order is k1-major
        ref = ref.reshape(n2_, n1_).T.reshape(-1)
        assert ret["exchanges"] == 2
    else:
        assert ret["exchanges"] == 3
    assert oracle.rel_l2(y[None], ref[None]) < 1e-12


def test_logical_ranks_match_oracle(oracle):
    """The in-process 'fake shard' driver (P logical ranks, block swaps instead of NCCL)."""
    for n, world in ((256, 4), (1536, 2), (4096, 8)):
        x = oracle.uniform_complex((n,), 3, np.complex128)
        plan = DistFFT1D(n, world, dtype=torch.complex128, ops=NumpyOps())
        per = n // world
        xs = [torch.from_numpy(x[r * per:(r + 1) * per].copy()) for r in range(world)]
        ys = plan.run_logical(xs)
        y = torch.cat(ys).numpy()
        assert oracle.rel_l2(y[None], oracle.fft(x[None])) < 1e-12, (n, world)
        back = torch.cat(plan.run_logical([t.clone() for t in ys], inverse=True)).numpy()
        assert oracle.rel_l2(back[None], (n * x)[None]) < 1e-12


def test_shard_range_and_factors():
    for total, world in ((65536, 8), (8192, 4), (10, 3), (7, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    assert choose_factors(1 << 30, 8) == (1 << 15, 1 << 15)
    assert choose_factors(1 << 20, 2) == (1 << 10, 1 << 10)
    n1, n2 = choose_factors(1 << 21, 4)
    assert n1 * n2 == 1 << 21 and n1 % 4 == 0 and n2 % 4 == 0
    with pytest.raises(ValueError):
        choose_factors(30, 4)
