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


class PeerBuffers:
    """Two exchange buffers per rank, mapped into every other rank with CUDA IPC so kernels can store straight
    into a peer's HBM over NVLink (one process per GPU).  No CPU fallback; needs world > 1 real CUDA ranks."""

    def __init__(self, elems: int, dtype, rank: int, world: int, group=None, device=None):
        import torch.distributed as dist

        self.lib = L.load()
        self.rank, self.world, self.group = rank, world, group
        self.bytes = elems * (8 if dtype == torch.complex64 else 16)
        self.local = []     # my two buffers (device pointers)
        self.tables = []    # per buffer: ctypes array of `world` device pointers (index = rank)
        self._imported = []
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        for _ in range(2):
            ptr = ctypes.c_void_p()
            L.check(self.lib.ssfft_malloc(ctypes.byref(ptr), self.bytes), "ssfft_malloc")
            handle = ctypes.create_string_buffer(64)
            L.check(self.lib.ssfft_ipc_export(ptr, handle), "ssfft_ipc_export")
            mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(device)
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine, group=group)
            table = (ctypes.c_void_p * world)()
            for r in range(world):
                if r == rank:
                    table[r] = ptr.value
                else:
                    peer = ctypes.c_void_p()
                    raw = bytes(gathered[r].cpu().numpy().tobytes())
                    L.check(self.lib.ssfft_ipc_import(raw, ctypes.byref(peer)), f"ssfft_ipc_import(rank {r})")
                    table[r] = peer.value
                    self._imported.append(peer)
            self.local.append(ptr)
            self.tables.append(table)
        self.barrier()

    def barrier(self):
        """Stream-ordered barrier: every rank's preceding kernels (incl. its peer stores) are complete after it."""
        import torch.distributed as dist

        dist.all_reduce(self._flag, group=self.group)

    def close(self):
        for p in self._imported:
            self.lib.ssfft_ipc_close(p)
        for p in self.local:
            self.lib.ssfft_free(p)
        self._imported, self.local = [], []


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
        self.profile = None  # set to {} to accumulate per-phase milliseconds (CUDA events) in _run

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
        prof = self.profile is not None and x_local.is_cuda
        marks = []

        def mark(name):
            if prof:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        send = self._pack_rows(x_local, w1)
        mark("transpose1")
        recv = self._all_to_all(send, w2)
        mark("a2a1")
        c = self.ops.permute102(recv.view(self.world, self.b, self.a), w1).view(self.b, self.n1)
        mark("permute1")
        f = self.ops.fft_rows(c, inverse, w2.view(self.b, self.n1))
        mark("fft_n1")
        send = self.ops.transpose(f, w1, row0=self.rank * self.b, n_total=self.n, inverse=inverse)
        send = send.view(self.world, self.a, self.b)
        mark("transpose2+twiddle")
        recv = self._all_to_all(send, w2)
        mark("a2a2")
        c = self.ops.permute102(recv.view(self.world, self.a, self.b), w1).view(self.a, self.n2)
        mark("permute2")
        g = self.ops.fft_rows(c, inverse, (out_local if self.transposed_output else w2).view(self.a, self.n2))
        mark("fft_n2")
        if not self.transposed_output:
            send = self.ops.transpose(g, w1).view(self.world, self.b, self.a)
            mark("transpose3")
            recv = self._all_to_all(send, w2)
            mark("a2a3")
            self._natural(recv, out_local)
            mark("permute3")
        if prof:
            torch.cuda.synchronize()
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                self.profile[name] = self.profile.get(name, 0.0) + e0.elapsed_time(e1)
            self.profile["calls"] = self.profile.get("calls", 0) + 1
        return out_local

    # ---- fused path: every exchange is ONE kernel that transposes (+ twiddles) and stores straight into the peers'
    # HBM over NVLink at the final position (ssfft_exchange_transpose); no pack, no NCCL all-to-all, no unpack.
    def enable_peer_exchange(self, device):
        """Allocate and IPC-share the exchange buffers (collective call).  Needs world > 1 CUDA ranks."""
        self._peers = PeerBuffers(self.n // self.world, self.dtype, self.rank, self.world, self.group, device)
        return self

    def _run_p2p(self, x_local, out_local, inverse):
        ops, pb = self.ops, self._peers
        lib, prec = ops.lib, ops.prec
        esz = 8 if self.dtype == torch.complex64 else 16
        per = self.n // self.world
        if self._work is None or self._work[0].device != x_local.device:
            self._work = (torch.empty(per, dtype=self.dtype, device=x_local.device),
                          torch.empty(per, dtype=self.dtype, device=x_local.device))
        w = self._work[0]
        stream = ops._stream(x_local)
        a, b, n1, n2, r = self.a, self.b, self.n1, self.n2, self.rank
        inv = 1 if inverse else 0
        direction = L.SSFFT_INVERSE if inverse else L.SSFFT_FORWARD
        prof = self.profile is not None
        marks = []

        def mark(name):
            if prof:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        def plan_of(n):
            pl = ops._plans.get(n)
            if pl is None:
                pl = ops._plans[n] = FFT(n, dtype="float32" if prec == L.SSFFT_F32 else "float64")
            return pl._plan

        bufA, bufB = pb.local[0], pb.local[1]
        # every rank has finished reading its exchange buffers in the previous call (its copy-out of buffer A) before
        # anyone stores into them again
        pb.barrier()
        mark("start")
        # exchange 1: x[a][n2] -> peers' A as [b][n1]  (my rows land at columns r*a ..)
        L.check(lib.ssfft_exchange_transpose(x_local.data_ptr(), pb.tables[0], self.world, a, n2, n1, r * a, 0, 0, 0, prec,
                                             stream), "ssfft_exchange_transpose")
        pb.barrier()
        mark("exchange1")
        L.check(lib.ssfft_exec_c2c(plan_of(n1), bufA, w.data_ptr(), b, direction, stream), "ssfft_exec_c2c")
        mark("fft_n1")
        # exchange 2 (+ twiddle W_N^(n2 k1)): [b][n1] -> peers' B as [a][n2]
        L.check(lib.ssfft_exchange_transpose(w.data_ptr(), pb.tables[1], self.world, b, n1, n2, r * b, r * b, self.n, inv,
                                             prec, stream), "ssfft_exchange_transpose")
        pb.barrier()
        mark("exchange2+twiddle")
        if self.transposed_output:
            L.check(lib.ssfft_exec_c2c(plan_of(n2), bufB, out_local.data_ptr(), a, direction, stream), "ssfft_exec_c2c")
            mark("fft_n2")
        else:
            L.check(lib.ssfft_exec_c2c(plan_of(n2), bufB, w.data_ptr(), a, direction, stream), "ssfft_exec_c2c")
            mark("fft_n2")
            # exchange 3: [a][k2] -> peers' A as [b][n1] == natural order X[k2*n1 + k1]
            L.check(lib.ssfft_exchange_transpose(w.data_ptr(), pb.tables[0], self.world, a, n2, n1, r * a, 0, 0, 0, prec,
                                                 stream), "ssfft_exchange_transpose")
            pb.barrier()
            mark("exchange3")
            L.check(lib.ssfft_memcpy_d2d(out_local.data_ptr(), bufA, per * esz, stream), "ssfft_memcpy_d2d")
            mark("copy_out")
        self.exchanges += 2 if self.transposed_output else 3
        if prof:
            torch.cuda.synchronize()
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                self.profile[name] = self.profile.get(name, 0.0) + e0.elapsed_time(e1)
            self.profile["calls"] = self.profile.get("calls", 0) + 1
        return out_local

    def fft(self, x_local: torch.Tensor, out_local: torch.Tensor) -> torch.Tensor:
        if getattr(self, "_peers", None) is not None:
            return self._run_p2p(x_local, out_local, False)
        return self._run(x_local, out_local, False)

    def ifft(self, x_local: torch.Tensor, out_local: torch.Tensor) -> torch.Tensor:
        if getattr(self, "_peers", None) is not None:
            return self._run_p2p(x_local, out_local, True)
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


class LocalDistFFT1D:
    """One length-N complex transform sharded over several GPUs of THIS process, through the C ABI
    (ssfft_dist_plan_create / ssfft_dist_exec_c2c, include/ssfft.h): exchanges by peer stores over NVLink, the exchange of
    a chunk under the FFT of the next one, the last exchange straight into the output shards.  `devices` may repeat an
    index (logical ranks on one GPU).  Shards: lists of CUDA tensors of N / len(devices) complex elements, one per device,
    block r of the natural order on devices[r]."""

    def __init__(self, n: int, devices, dtype=torch.complex64, transposed_output: bool = False):
        self.lib = L.load()
        self.n, self.devices = n, list(devices)
        self.dtype = dtype
        prec = L.SSFFT_F32 if dtype == torch.complex64 else L.SSFFT_F64
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        self._plan = ctypes.c_void_p()
        L.check(self.lib.ssfft_dist_plan_create(ctypes.byref(self._plan), prec, n, len(self.devices), arr,
                                                1 if transposed_output else 0), f"ssfft_dist_plan_create(n={n})")
        self.n1 = self.lib.ssfft_dist_plan_factor(self._plan, 0)
        self.n2 = self.lib.ssfft_dist_plan_factor(self._plan, 1)

    def describe(self) -> str:
        buf = ctypes.create_string_buffer(1024)
        self.lib.ssfft_dist_plan_describe(self._plan, buf, 1024)
        return buf.value.decode()

    def _run(self, xs, outs, direction):
        P = len(self.devices)
        assert len(xs) == P and len(outs) == P
        per = self.n // P
        for r, (x, o) in enumerate(zip(xs, outs)):
            assert x.is_cuda and o.is_cuda and x.device.index == self.devices[r] and o.device.index == self.devices[r]
            assert x.dtype == self.dtype and o.dtype == self.dtype and x.numel() == per and o.numel() == per
            assert x.is_contiguous() and o.is_contiguous()
        tin = (ctypes.c_void_p * P)(*[x.data_ptr() for x in xs])
        tout = (ctypes.c_void_p * P)(*[o.data_ptr() for o in outs])
        L.check(self.lib.ssfft_dist_exec_c2c(self._plan, tin, tout, direction), "ssfft_dist_exec_c2c")

    def fft(self, xs, outs):
        self._run(xs, outs, L.SSFFT_FORWARD)

    def ifft(self, xs, outs):
        self._run(xs, outs, L.SSFFT_INVERSE)

    def synchronize(self):
        L.check(self.lib.ssfft_dist_synchronize(self._plan), "ssfft_dist_synchronize")

    def close(self):
        if self._plan:
            self.lib.ssfft_dist_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
