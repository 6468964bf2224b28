"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), distributed four-step vs the oracle."""
import os, sys, math
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
from fft_b200.dist import DistFFT1D
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
for n in (1 << 16, 1 << 20, 3 * (1 << 14)):
    x = O.uniform_complex((n,), 21, np.complex64)
    per = n // world
    xl = torch.from_numpy(x[rank * per:(rank + 1) * per].copy()).to(dev)
    outs = {}
    for mode in ("nccl", "p2p"):
        plan = DistFFT1D(n, world, rank, dtype=torch.complex64)
        if mode == "p2p":
            plan.enable_peer_exchange(dev)
        y = torch.empty_like(xl)
        plan.fft(xl, y)
        back = torch.empty_like(xl)
        plan.ifft(y, back)
        g = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(g, y)
        gb = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(gb, back)
        outs[mode] = (torch.cat(g).cpu().numpy(), torch.cat(gb).cpu().numpy())
        if mode == "p2p":
            plan._peers.close()
    if rank == 0:
        ref = O.run(O.KIND_C2C_FWD, x[None], n, threads=2)[0]
        lim = 1e-6 * math.log2(n)
        for mode, (y, back) in outs.items():
            assert O.rel_l2(y[None], ref) <= lim, (mode, n, O.rel_l2(y[None], ref))
            assert O.rel_l2(back[None], (n * x)[None]) <= 2 * lim, (mode, n)
        assert O.rel_l2(outs["p2p"][0][None], outs["nccl"][0][None]) <= 1e-6, n
dist.barrier()
if rank == 0:
    print("MULTI-GPU-OK")
dist.destroy_process_group()
