"""Multi-GPU execution of the FFT hot path.

Two cases, one process per GPU (torch.distributed over NCCL for the plumbing):

* batched transforms shard by batch -- ``shard_range`` -- with NO collective on the data path;
* ONE huge 1-D transform (BASELINE config 5, N = 2^30) is sharded with the four-step / transpose algorithm:
  N = N1 * N2, natural-order input and output, block distributed (rank r owns elements
  [r*N/P, (r+1)*N/P)).  Three exchange steps, each ``transpose -> all-to-all -> block permute``:

      x[n1][n2] rows        --T,a2a,P-->  [n2 local][n1]   --FFT_N1, twiddle W_N^(n2 k1)-->
      [n2 local][k1]        --T,a2a,P-->  [k1 local][n2]   --FFT_N2-->
      [k1 local][k2]        --T,a2a,P-->  [k2 local][k1] = X[k2*N1 + k1]   (natural order)

  ``transposed_output=True`` skips the third exchange and returns X[k1 + N1*k2] for the local k1 rows.
  The local steps are this library's kernels through the C ABI (batched contiguous FFTs,
  ssfft_transpose_twiddle, ssfft_permute102); the exchange is ``all_to_all_single`` (NCCL over NVLink).

The local operations are injected (``ops``) so the rank/index logic can be tested on CPU with gloo and a
numpy stand-in (tests/test_dist_gloo.py); the default ``CudaOps`` has no CPU fallback.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib as L
from .api import FFT


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous batch shard [begin, end) of `total` independent transforms for `rank` of `world`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return total * rank // world, total * (rank + 1) // world


def choose_factors(n: int, world: int, max_len: int = 1 << 20) -> tuple[int, int]:
    """N = N1 * N2 with both divisible by `world`, as balanced as possible, each <= max_len."""
    best = None
    d = 1
    while d * d <= n:
        if n % d == 0:
            for n1 in (d, n // d):
                n2 = n // n1
                if n1 % world == 0 and n2 % world == 0 and n1 <= max_len and n2 <= max_len:
                    score = abs(math.log(n1 / n2)) + (1e-9 if n1 > n2 else 0.0)  # ties: n1 <= n2
                    if best is None or score < best[0]:
                        best = (score, n1, n2)
        d += 1
    if best is None:
        raise ValueError(f"cannot split N={n} into two factors divisible by world={world}")
    return best[1], best[2]


class CudaOps:
    """Local steps on the GPU through the C ABI (no CPU fallback)."""

    def __init__(self, dtype=torch.complex64):
        self.lib = L.load()
        self.dtype = dtype
        self.prec = L.SSFFT_F32 if dtype == torch.complex64 else L.SSFFT_F64
        self._plans = {}

    def _stream(self, t):
        return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    def fft_rows(self, x: torch.Tensor, inverse: bool, out: torch.Tensor) -> torch.Tensor:
        n = x.shape[-1]
        plan = self._plans.get(n)
        if plan is None:
            plan = self._plans[n] = FFT(n, dtype="float32" if self.prec == L.SSFFT_F32 else "float64")
        (plan.ifft if inverse else plan.fft)(x, out)
        return out

    def transpose(self, x: torch.Tensor, out: torch.Tensor, row0: int = 0, n_total: int = 0,
                  inverse: bool = False) -> torch.Tensor:
        rows, cols = x.shape
        L.check(self.lib.ssfft_transpose_twiddle(x.data_ptr(), out.data_ptr(), 1, rows, cols, row0, n_total,
                                                 1 if inverse else 0, self.prec, self._stream(x)),
                "ssfft_transpose_twiddle")
        return out.view(cols, rows)

    def permute102(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        a, b, run = x.shape
        L.check(self.lib.ssfft_permute102(x.data_ptr(), out.data_ptr(), a, b, run, self.prec, self._stream(x)),
                "ssfft_permute102")
        return out.view(b, a, run)


class DistFFT1D:
    """One length-N complex transform sharded over `world` ranks (natural order in, natural order out)."""

    def __init__(self, n: int, world: int, rank: int | None = None, group=None, dtype=torch.complex64, ops=None,
                 n1: int | None = None, transposed_output: bool = False):
        self.n, self.world, self.rank, self.group = n, world, rank, group
        self.dtype = dtype
        if n1 is None:
            n1, n2 = choose_factors(n, world)
        else:
            n2 = n // n1
        if n1 * n2 != n or n1 % world or n2 % world:
            raise ValueError(f"N={n} = {n1} x {n2} is not divisible by world={world} in both factors")
        self.n1, self.n2 = n1, n2
        self.a, self.b = n1 // world, n2 // world  # rows of x / columns owned per rank
        self.ops = ops if ops is not None else CudaOps(dtype)
        self.transposed_output = transposed_output
        self.exchanges = 0
        self._work = None  # two work buffers of N/P elements, allocated on first use and kept

    # ---- the three local phases; each returns the packed send buffer [P][..][..] for the next exchange
    def _pack_rows(self, x_local, work):
        """[a][n2] -> transpose -> [n2][a] == [P][b][a] (destination-major)."""
        return self.ops.transpose(x_local.view(self.a, self.n2), work).view(self.world, self.b, self.a)

    def _columns(self, recv, rank, inverse, work_a, work_b):
        """recv [P][b][a] -> [b][n1] -> FFT over n1 -> twiddle + transpose -> [n1][b] == [P][a][b]."""
        c = self.ops.permute102(recv.view(self.world, self.b, self.a), work_a).view(self.b, self.n1)
        f = self.ops.fft_rows(c, inverse, work_b.view(self.b, self.n1))
        t = self.ops.transpose(f, work_a, row0=rank * self.b, n_total=self.n, inverse=inverse)
        return t.view(self.world, self.a, self.b)

    def _rows(self, recv, inverse, work_a, work_b):
        """recv [P][a][b] -> [a][n2] -> FFT over n2 -> [a][k2] (== X[k1 + N1*k2] for the local k1)."""
        c = self.ops.permute102(recv.view(self.world, self.a, self.b), work_a).view(self.a, self.n2)
        return self.ops.fft_rows(c, inverse, work_b.view(self.a, self.n2))

    def _natural(self, recv, out):
        """recv [P][b][a] -> [b][n1]: natural-order block of X."""
        return self.ops.permute102(recv.view(self.world, self.b, self.a), out).view(-1)

    # ---- exchange: real (torch.distributed) or in-process over a list of logical ranks
    def _all_to_all(self, send: torch.Tensor, recv: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        self.exchanges += 1
        dist.all_to_all_single(torch.view_as_real(recv.view(-1)), torch.view_as_real(send.contiguous().view(-1)),
                               group=self.group)
        return recv

    def _run(self, x_local, out_local, inverse):
        if self.rank is None:
            raise ValueError("rank is required for the distributed call; use run_logical() for in-process shards")
        per = self.n // self.world
        if self._work is None or self._work[0].device != x_local.device:
            self._work = (torch.empty(per, dtype=self.dtype, device=x_local.device),
                          torch.empty(per, dtype=self.dtype, device=x_local.device))
        w1, w2 = self._work
        send = self._pack_rows(x_local, w1)
        recv = self._all_to_all(send, w2)
        send = self._columns(recv, self.rank, inverse, w1, w2)          # result lives in w1
        recv = self._all_to_all(send, w2)
        g = self._rows(recv, inverse, w1, out_local if self.transposed_output else w2)
        if self.transposed_output:
            return out_local
        send = self.ops.transpose(g, w1).view(self.world, self.b, self.a)
        recv = self._all_to_all(send, w2)
        self._natural(recv, out_local)
        return out_local

    def fft(self, x_local: torch.Tensor, out_local: torch.Tensor) -> torch.Tensor:
        return self._run(x_local, out_local, False)

    def ifft(self, x_local: torch.Tensor, out_local: torch.Tensor) -> torch.Tensor:
        return self._run(x_local, out_local, True)

    # ---- "fake shard" mode: all logical ranks in one process (single-GPU unit test of the same code path)
    def run_logical(self, xs: list, inverse: bool = False) -> list:
        P, per = self.world, self.n // self.world
        assert len(xs) == P

        def new():
            return [torch.empty(per, dtype=self.dtype, device=xs[0].device) for _ in range(P)]

        def swap(sends):  # all-to-all among the logical ranks: block j of rank i -> block i of rank j
            self.exchanges += 1
            recvs = new()
            chunk = per // P
            for i in range(P):
                for j in range(P):
                    recvs[j].view(-1)[i * chunk:(i + 1) * chunk] = sends[i].reshape(-1)[j * chunk:(j + 1) * chunk]
            return recvs

        w1, w2 = new(), new()
        sends = [self._pack_rows(xs[r], w1[r]) for r in range(P)]
        recvs = swap(sends)
        sends = [self._columns(recvs[r], r, inverse, w1[r], w2[r]) for r in range(P)]
        recvs = swap(sends)
        w3 = new()
        gs = [self._rows(recvs[r], inverse, w1[r], w3[r]) for r in range(P)]
        if self.transposed_output:
            return [g.reshape(-1) for g in gs]
        sends = [self.ops.transpose(gs[r], w1[r]).view(P, self.b, self.a) for r in range(P)]
        recvs = swap(sends)
        outs = new()
        return [self._natural(recvs[r], outs[r]) for r in range(P)]
