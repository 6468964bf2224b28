#!/usr/bin/env python
"""bench.py -- headline benchmark of the FFT hot path (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref)

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE
config 2: FFT<float> C2C, N=4096 x 65536 transforms per GPU (4.29 GB of algorithmic HBM traffic per step,
34x the 126 MB L2, so no L2 flush is needed between steps).  Multi-GPU runs shard by batch (weak scaling:
every rank transforms its own 65536 x 4096 batch; there is no collective on the data path).

  value     whole-job GFLOP/s (5 N log2 N per transform), inputs resident in HBM, CUDA-event timed
  e2e       same metric through the host-pointer call a reference user makes (ssfft_exec_host):
            pinned host buffers, H2D + kernels + D2H all inside the timed region
  roofline  algorithmic bytes (2 * N * sizeof(complex) per transform) / kernel time vs measured HBM peak
  cpu_baseline  the reference's CPU implementation on this box's host cores (reported, not a target)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Batched fp32 C2C FFT GFLOP/s (5N·log2N) & % of HBM roofline, N=4096"
SEED = 20261017


def metric_for(workload):
    """BASELINE.json's metric is quoted on config 2 (the default workload); other workloads name their own size."""
    kind, dtype, n, _ = WORKLOADS[workload]
    if workload == "c2":
        return METRIC
    p = "fp32" if dtype == "float32" else "fp64"
    if kind == "c2c":
        return f"Batched {p} C2C FFT GFLOP/s (5N·log2N) & % of HBM roofline, N={n}"
    return f"Batched {p} RealFFT forward+inverse GFLOP/s (2 x 2.5N·log2N) & % of HBM roofline, N={n}"

WORKLOADS = {
    # name: (kind, dtype, N, batch)
    "c2": ("c2c", "float32", 4096, 65536),
    "c3": ("real", "float32", 65536, 8192),
    "c4-1000": ("c2c", "float32", 1000, 65536),
    "c4-2187": ("c2c", "float32", 2187, 65536),
    "c4-3125": ("c2c", "float32", 3125, 65536),
    "c4-6000": ("c2c", "float32", 6000, 65536),
    "c4-1000-f64": ("c2c", "float64", 1000, 65536),
    "c4-2187-f64": ("c2c", "float64", 2187, 65536),
    "c4-3125-f64": ("c2c", "float64", 3125, 65536),
    "c4-6000-f64": ("c2c", "float64", 6000, 65536),
    "c1": ("c2c", "float64", 1024, 1),
    # one transform of 2^30 points sharded over all ranks (distributed four-step, strong scaling)
    "c5": ("dist", "float32", 1 << 30, 1),
    "c5-small": ("dist", "float32", 1 << 24, 1),
}


def flops_per_transform(kind, n):
    # complex: 5 N log2 N; real: 2.5 N log2 N per direction, forward + inverse are both run
    return 5.0 * n * math.log2(n) if kind == "c2c" else 2 * 2.5 * n * math.log2(n)


def alg_bytes_per_transform(kind, n, dtype):
    sz = 4 if dtype == "float32" else 8
    # C2C: read N complex + write N complex.  Real fwd+inv: 2 * (N reals + N/2 complex) = 4 N reals
    return 2 * n * 2 * sz if kind == "c2c" else 4 * n * sz


def bind_to_gpu_numa_node(index):
    """Run this process on the CPUs next to GPU `index` (its PCIe root complex) before pinned host memory is allocated:
    pages are placed on the node of the thread that first touches them, and a host buffer on the far socket halves the
    copy rate of every rank that shares the inter-socket link.  Returns a short description for the JSON line."""
    try:
        import torch

        prop = torch.cuda.get_device_properties(index)
        bus = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{getattr(prop, 'pci_device_id', 0):02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        with open(base + "/local_cpulist") as f:
            cpulist = f.read().strip()
        node = open(base + "/numa_node").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": bus, "numa_node": node, "cpus": len(cpus) or len(allowed)}
    except Exception as exc:  # containers without sysfs access: leave the affinity alone
        return {"error": repr(exc)[:120]}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def profiled_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` summary of the same command (profiles/*_traffic.json, made by tools/summarize_ncu.py)."""
    files = {"c2": "c2_fused4096_tma_r02_traffic.json"}
    name = files.get(workload)
    if not name:
        return None, None
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            data = json.load(f)
        kernel, launches = next(iter(data.items()))
        l0 = launches[0]
        return l0["dram_read_bytes"] + l0["dram_write_bytes"], f"profiles/{name}"
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import oracle as O  # the one place bench.py may execute oracle/

    kind, dtype, n, batch = WORKLOADS[args.workload]
    impl = "reference" if O.have_reference() else "port"
    cores = os.cpu_count() or 1
    rdt = np.float32 if dtype == "float32" else np.float64
    # bounded sample: ~15 CPU-seconds of transforms per step, never more than the workload itself
    probe = O.uniform(2 * n * 64 if kind == "c2c" else n * 64, SEED, rdt)
    probe = probe.view(np.complex64 if dtype == "float32" else np.complex128).reshape(64, n) if kind == "c2c" \
        else probe.reshape(64, n)
    k_fwd = O.KIND_C2C_FWD if kind == "c2c" else O.KIND_R2C
    t0 = time.perf_counter()
    O.run(k_fwd, probe, n, 1, impl)
    per = max((time.perf_counter() - t0) / 64, 1e-7)
    sample = int(min(batch, max(cores * 8, 15.0 / per / max(args.steps + args.warmup, 1))))
    x = O.uniform((2 if kind == "c2c" else 1) * n * sample, SEED, rdt)
    x = x.view(np.complex64 if dtype == "float32" else np.complex128).reshape(sample, n) if kind == "c2c" \
        else x.reshape(sample, n)

    def step():
        if kind == "c2c":
            return O.run(O.KIND_C2C_FWD, x, n, cores, impl)[1]
        y, s1 = O.run(O.KIND_R2C, x, n, cores, impl)
        _, s2 = O.run(O.KIND_C2R, y, n, cores, impl)
        return s1 + s2

    for _ in range(args.warmup):
        step()
    secs = [step() for _ in range(args.steps)]
    total = sum(secs)
    value = sample * args.steps * flops_per_transform(kind, n) / total / 1e9
    line = {
        "impl": "reference", "metric": metric_for(args.workload), "value": value, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == "float32" else "f64",
        "data": "synthetic uniform[-0.5,0.5), counter-based generator, seed %d" % SEED,
        # the same workload string as the GPU arm prints (the driver compares them); what a step really covers is a
        # bounded sample of that workload, stated next to it and in cpu_baseline.sample
        "config": {"workload": f"{args.workload}: {kind} {dtype} N={n} x {batch} transforms per GPU"
                               + (" (forward + inverse per step)" if kind != "c2c" else " (forward)"),
                   "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": cores, "kind": impl,
                         "sample": f"{sample} transforms of N={n} per step, batch split over {cores} host threads, "
                                   "one FFT object per thread, plan build excluded"},
        "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_dist(args, n, rank, world, dev, dist):
    """BASELINE config 5: ONE length-n complex transform sharded over all ranks (fft_b200/dist.py)."""
    line = measure_dist(args, n, rank, world, dev, dist, steps=args.steps, warmup=args.warmup, verify=args.verify)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def measure_dist(args, n, rank, world, dev, dist, steps, warmup, verify):
    import torch

    import fft_b200
    from fft_b200.dist import DistFFT1D

    per = n // world
    plan = DistFFT1D(n, world, rank, dtype=torch.complex64)
    x = torch.empty(per, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    fft_b200.fill_uniform(x, SEED, first_idx=rank * per * 2)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    exchange = "all_to_all_single over NCCL"
    if world == 1:
        plan._all_to_all = lambda send, recv: recv.copy_(send.reshape(-1))  # P = 1: the exchange is the identity
    elif not args.nccl_exchange:
        plan.enable_peer_exchange(dev)
        exchange = "fused transpose + direct peer stores over NVLink (CUDA IPC), no NCCL on the data path"
    for _ in range(warmup):
        plan.fft(x, y)
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = fft_b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        plan.fft(x, y)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = fft_b200.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / steps
    # per-phase breakdown (outside the timed region): two extra calls with CUDA events between the phases
    plan.profile = {}
    for _ in range(2):
        plan.fft(x, y)
    phases = {k: round(v / plan.profile["calls"], 3) for k, v in plan.profile.items() if k != "calls"}
    plan.profile = None
    flops = 5.0 * n * math.log2(n)
    peak, peak_src = measured_peak()
    alg_bytes_gpu = 2 * per * 8
    exch_bytes_gpu = 3 * per * 8 * (world - 1) / max(world, 1)  # sent per GPU per step over NVLink
    # size-independent checks (the oracle cannot run 2^30 casually): single tone -> N delta,
    # Parseval, and ifft(fft(x)) = N x on the random input
    checks = None
    if verify:
        def allsum(v):
            t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t)
            return float(t.item())
        f0 = 123456789 % n
        idx = torch.arange(per, device=dev, dtype=torch.int64) + rank * per
        ph = ((idx * f0) % n).double() * (2.0 * math.pi / n)
        tone = torch.complex(torch.cos(ph), torch.sin(ph)).to(torch.complex64)
        del idx, ph
        yt = torch.empty_like(tone)
        plan.fft(tone, yt)
        e_all = allsum((yt.real.double() ** 2 + yt.imag.double() ** 2).sum().item())
        own = f0 // per
        peak_v = yt[f0 - own * per].item() if rank == own else 0j
        pr, pi = allsum(peak_v.real), allsum(peak_v.imag)
        tone_err = math.sqrt(max(e_all - (pr * pr + pi * pi), 0.0) + (pr - n) ** 2 + pi ** 2) / n
        del tone, yt
        ex = allsum((x.real.double() ** 2 + x.imag.double() ** 2).sum().item())
        ey = allsum((y.real.double() ** 2 + y.imag.double() ** 2).sum().item())
        back = torch.empty_like(x)
        plan.ifft(y, back)
        num = allsum(((back.real.double() - n * x.real.double()) ** 2 + (back.imag.double() - n * x.imag.double()) ** 2).sum().item())
        rt_err = math.sqrt(num / (n * n * ex))
        lim = 1e-6 * math.log2(n)
        checks = {"tone_rel_err": tone_err, "parseval_rel_err": abs(ey / (n * ex) - 1.0), "roundtrip_rel_l2": rt_err,
                  "tolerance": lim, "ok": bool(tone_err <= lim and rt_err <= 2 * lim and abs(ey / (n * ex) - 1.0) < 1e-4)}
        del back
    line = None
    if rank == 0:
        line = {
            "metric": "Distributed fp32 C2C FFT GFLOP/s (5N·log2N), single transform N=%d" % n,
            "value": flops / (ms_step * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32",
            "data": f"synthetic uniform[-0.5,0.5), counter-based generator, seed {SEED}",
            "config": {"workload": f"{args.workload}: one c2c float32 transform of N={n} = {plan.n1} x {plan.n2}, "
                                   f"block-sharded over {world} GPUs, natural order in and out",
                       "parallelism": f"four-step x{world}, 3 exchanges: {exchange}",
                       "phase_ms_rank0": phases},
            "roofline": {"bound": "hbm", "achieved": alg_bytes_gpu / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_bytes_gpu / (ms_step * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "nvlink_bytes_per_gpu_per_step": exch_bytes_gpu,
                         "nvlink_GBps_per_gpu_if_exchange_only": exch_bytes_gpu / (ms_step * 1e-3) / 1e9},
            "clocks": clocks, "gpu_launches": launches,
        }
        if checks:
            line["checks"] = checks
    barrier()
    if getattr(plan, "_peers", None) is not None:  # unmap the peers' exchange buffers, free mine
        plan._peers.close()
    del plan, x, y
    torch.cuda.empty_cache()
    return line


def measure_dist_local(n, world, iters=3):
    """BASELINE config 5 through ssfft_dist_plan_create / ssfft_dist_exec_c2c: ONE process, `world` GPUs, natural order."""
    import torch

    import fft_b200
    from fft_b200.dist import LocalDistFFT1D

    per = n // world
    out = {}
    for transposed in (False, True):
        plan = LocalDistFFT1D(n, list(range(world)), transposed_output=transposed)
        xs, ys = [], []
        for r in range(world):
            with torch.cuda.device(r):
                x = torch.empty(per, dtype=torch.complex64, device=f"cuda:{r}")
                fft_b200.fill_uniform(x, SEED, first_idx=r * per * 2)
                xs.append(x)
                ys.append(torch.empty_like(x))
        for r in range(world):
            torch.cuda.synchronize(r)
        for _ in range(2):
            plan.fft(xs, ys)
        plan.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            plan.fft(xs, ys)
        plan.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / iters
        key = "transposed_output" if transposed else "natural_order"
        out[key] = {"ms": ms, "gflops": 5.0 * n * math.log2(n) / ms / 1e6}
        if not transposed:
            ex = sum((x.real.double() ** 2 + x.imag.double() ** 2).sum().item() for x in xs)
            ey = sum((y.real.double() ** 2 + y.imag.double() ** 2).sum().item() for y in ys)
            backs = [torch.empty_like(x) for x in xs]
            plan.ifft(ys, backs)
            plan.synchronize()
            num = sum(((b.real.double() - n * x.real.double()) ** 2 + (b.imag.double() - n * x.imag.double()) ** 2).sum().item()
                      for b, x in zip(backs, xs))
            lim = 1e-6 * math.log2(n)
            rt = math.sqrt(num / (n * n * ex))
            out["checks"] = {"parseval_rel_err": abs(ey / (n * ex) - 1.0), "roundtrip_rel_l2": rt, "tolerance": lim,
                             "ok": bool(rt <= 2 * lim and abs(ey / (n * ex) - 1.0) < 1e-4)}
            out["plan"] = plan.describe()
            del backs
        plan.close()
        del xs, ys
        torch.cuda.empty_cache()
    out["workload"] = f"one c2c float32 transform of N={n} over {world} GPUs of one process (C ABI ssfft_dist_*), wall clock"
    return out


def measure_batched(name, dev, rank, world, dist, steps=5, warmup=3, subset=4):
    """One BASELINE config other than the headline: device-resident timing (CUDA events, max over ranks) plus the
    relative L2 error of a few transforms against the oracle in the same precision."""
    import numpy as np
    import torch

    import fft_b200

    kind, dtype, n, batch = WORKLOADS[name]
    tdt = torch.float32 if dtype == "float32" else torch.float64
    cdt = torch.complex64 if dtype == "float32" else torch.complex128
    if kind == "c2c":
        plan = fft_b200.FFT(n, dtype=dtype)
        x = torch.empty((batch, n), dtype=cdt, device=dev)
        y = torch.empty_like(x)
        fft_b200.fill_uniform(x, SEED, first_idx=rank * batch * n * 2)

        def step():
            plan.fft(x, y)
    else:
        plan = fft_b200.RealFFT(n, dtype=dtype)
        x = torch.empty((batch, n), dtype=tdt, device=dev)
        y = torch.empty((batch, n // 2), dtype=cdt, device=dev)
        z = torch.empty_like(x)
        fft_b200.fill_uniform(x, SEED, first_idx=rank * batch * n)

        def step():
            plan.fft(x, y)
            plan.ifft(y, z)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    peak, _ = measured_peak()
    out = {"workload": f"{kind} {dtype} N={n} x {batch} per GPU" + (" fwd+inv" if kind != "c2c" else ""), "ms": ms,
           "gflops": world * batch * flops_per_transform(kind, n) / (ms * 1e-3) / 1e9,
           "roofline_frac": batch * alg_bytes_per_transform(kind, n, dtype) / (ms * 1e-3) / 1e9 / peak,
           "plan": plan.describe()[:160]}
    if rank == 0:
        from oracle import oracle as O  # checker only

        xs = x[:subset].cpu().numpy()
        if kind == "c2c":
            ref = O.run(O.KIND_C2C_FWD, xs, n, subset)[0]
            out["parity_relL2_vs_oracle_on_a_subset"] = float(O.rel_l2(y[:subset].cpu().numpy(), ref))
        else:
            ref = O.rfft(xs)
            out["parity_relL2_vs_oracle_on_a_subset"] = float(O.rel_l2(y[:subset].cpu().numpy(), ref))
            back = O.run(O.KIND_C2R, ref, n, subset)[0]
            out["parity_relL2_inverse"] = float(O.rel_l2(z[:subset].cpu().numpy(), back))
        lim = (1e-6 if dtype == "float32" else 1e-14) * math.log2(n)
        out["parity_tolerance"] = lim
        out["parity_ok"] = bool(out["parity_relL2_vs_oracle_on_a_subset"] <= lim and out.get("parity_relL2_inverse", 0.0) <= 2 * lim)
    return out


def measure_c1(dev):
    """BASELINE config 1: ONE FFT<double> N=1024 transform, fft + ifft round trip -- device-resident and through the
    host call a reference user makes (numpy in / out -> ssfft_exec_host), beside the reference on one host core."""
    import numpy as np
    import torch

    import fft_b200
    from oracle import oracle as O  # checker / baseline only

    n = 1024
    plan = fft_b200.FFT(n, dtype="float64")
    xh = O.uniform_complex((1, n), SEED, np.complex128)
    x = torch.from_numpy(xh).to(dev)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    for _ in range(20):
        plan.fft(x, y)
        plan.ifft(y, z)
    torch.cuda.synchronize()
    iters = 300
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(iters):
        plan.fft(x, y)
        plan.ifft(y, z)
    ev1.record()
    torch.cuda.synchronize()
    dev_us = 1e3 * ev0.elapsed_time(ev1) / iters
    yh, zh = np.empty_like(xh), np.empty_like(xh)
    for _ in range(20):
        plan.fft(xh, yh)
        plan.ifft(yh, zh)
    t0 = time.perf_counter()
    for _ in range(iters):
        plan.fft(xh, yh)
        plan.ifft(yh, zh)
    host_us = 1e6 * (time.perf_counter() - t0) / iters
    impl = "reference" if O.have_reference() else "port"
    reps = np.repeat(xh, 2000, axis=0)
    secs = O.run(O.KIND_C2C_FWD, reps, n, 1, impl)[1] + O.run(O.KIND_C2C_INV, reps, n, 1, impl)[1]
    ref = O.fft(xh)
    err = float(O.rel_l2(yh, ref))
    rt = float(O.rel_l2(zh / n, xh))
    flops = 2 * 5.0 * n * math.log2(n)
    return {"workload": "c2c float64 N=1024 x 1, fft + ifft round trip", "device_resident_us_per_pair": dev_us,
            "e2e_host_call_us_per_pair": host_us, "e2e_gflops": flops / host_us / 1e3,
            "reference_cpu_us_per_pair_1_core": 1e6 * secs / 2000, "reference_kind": impl,
            "parity_relL2_vs_oracle_on_a_subset": err, "roundtrip_relL2": rt, "parity_tolerance": 1e-14 * math.log2(n),
            "parity_ok": bool(err <= 1e-14 * math.log2(n)),
            "note": "launch-latency bound: one 16 KiB transform cannot fill a GPU; the batched configs are the product"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override transforms per GPU (tests only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--verify", action="store_true", help="c5: tone / Parseval / round-trip checks at full size")
    ap.add_argument("--nccl-exchange", action="store_true", help="c5: use NCCL all_to_all instead of fused peer stores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import fft_b200

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    kind, dtype, n, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    tdt = torch.float32 if dtype == "float32" else torch.float64
    cdt = torch.complex64 if dtype == "float32" else torch.complex128
    esz = 4 if dtype == "float32" else 8

    # ---- inputs resident in HBM (each rank its own slice of the global synthetic stream)
    if kind == "dist":
        return run_dist(args, n, rank, world, dev, dist)
    if kind == "c2c":
        plan = fft_b200.FFT(n, dtype=dtype)
        x = torch.empty((batch, n), dtype=cdt, device=dev)
        y = torch.empty_like(x)
        fft_b200.fill_uniform(x, SEED, first_idx=rank * batch * n * 2)

        def step():
            plan.fft(x, y)
    else:
        plan = fft_b200.RealFFT(n, dtype=dtype)
        x = torch.empty((batch, n), dtype=tdt, device=dev)
        y = torch.empty((batch, n // 2), dtype=cdt, device=dev)
        z = torch.empty_like(x)
        fft_b200.fill_uniform(x, SEED, first_idx=rank * batch * n)

        def step():
            plan.fft(x, y)
            plan.ifft(y, z)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = fft_b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = fft_b200.launch_count() - launches0
    # keep the GPU busy a little longer so the clock sampler sees load even for short runs
    if rank == 0:
        t_end = time.time() + 0.6
        while time.time() < t_end:
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    flops_step = world * batch * flops_per_transform(kind, n)
    value = flops_step / (ms_step * 1e-3) / 1e9
    bytes_step_gpu = batch * alg_bytes_per_transform(kind, n, dtype)
    peak, peak_src = measured_peak()
    achieved = bytes_step_gpu / (ms_step * 1e-3) / 1e9  # per GPU

    # ---- end to end through the host-pointer API (pinned host memory, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        numa = bind_to_gpu_numa_node(local_rank)
        e2e_steps = max(1, min(args.steps, 8))
        host_in = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
        host_in.copy_(x)
        host_out = torch.empty(y.shape if kind == "c2c" else x.shape, dtype=(cdt if kind == "c2c" else tdt), pin_memory=True)
        host_mid = torch.empty(y.shape, dtype=cdt, pin_memory=True) if kind != "c2c" else None
        torch.cuda.synchronize()
        hin, hout = host_in.numpy(), host_out.numpy()
        hmid = host_mid.numpy() if host_mid is not None else None

        def e2e_step():
            if kind == "c2c":
                plan.fft(hin, hout)
            else:
                plan.fft(hin, hmid)
                plan.ifft(hmid, hout)

        e2e_step()  # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        io = batch * n * 2 * esz if kind == "c2c" else batch * n * esz
        e2e = {"value": flops_step * e2e_steps / el / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": io if kind == "c2c" else 2 * io, "d2h_bytes_per_step": io if kind == "c2c" else 2 * io,
               "ms_per_step": 1e3 * el / e2e_steps, "steps": e2e_steps,
               "api": "fft_b200.FFT.fft(host_in, host_out) -> ssfft_exec_host (pinned buffers, sliced H2D/compute/D2H overlap)",
               "host_numa": numa}
        # sanity: the e2e result equals the device-resident result
        if kind == "c2c":
            assert torch.equal(host_out[:4], y[:4].cpu()), "e2e output differs from device output"
        # what the link itself can do: the same bytes as plain pinned copies, both directions at once, no kernels
        try:
            cs_in, cs_out = torch.cuda.Stream(), torch.cuda.Stream()
            dbuf_in = torch.empty(x.shape, dtype=x.dtype, device=dev)
            dbuf_out = torch.empty(host_out.shape, dtype=host_out.dtype, device=dev)
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                with torch.cuda.stream(cs_in):
                    dbuf_in.copy_(host_in, non_blocking=True)
                with torch.cuda.stream(cs_out):
                    host_out.copy_(dbuf_out, non_blocking=True)
            torch.cuda.synchronize()
            el_c = (time.perf_counter() - t0) / 2
            if world > 1:
                t = torch.tensor([el_c], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                el_c = float(t.item())
            e2e["copy_only_ms_per_step"] = 1e3 * el_c
            e2e["ceiling_GBps_per_direction_per_gpu"] = e2e["h2d_bytes_per_step"] / el_c / 1e9
            e2e["frac_of_copy_ceiling"] = el_c / (el / e2e_steps)
            del dbuf_in, dbuf_out
        except Exception as exc:
            e2e["ceiling_error"] = repr(exc)[:200]
        del host_in, host_out

    # ---- CPU baseline: the reference's own implementation on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O  # checker / baseline only

        impl = "reference" if O.have_reference() else "port"
        cores = os.cpu_count() or 1
        per_transform_s = 60e-6 * (n * math.log2(n)) / (4096 * 12) * (20 if any(n % p == 0 for p in (5, 7, 11, 13)) else 1)
        sample = int(min(batch, max(cores * 4, 15.0 / per_transform_s)))
        xs = x[:sample].cpu().numpy()
        k = O.KIND_C2C_FWD if kind == "c2c" else O.KIND_R2C
        O.run(k, xs[: max(cores, 1)], n, cores, impl)
        best = None
        for _ in range(2):
            if kind == "c2c":
                secs = O.run(k, xs, n, cores, impl)[1]
            else:
                ys, s1 = O.run(O.KIND_R2C, xs, n, cores, impl)
                secs = s1 + O.run(O.KIND_C2R, ys, n, cores, impl)[1]
            best = secs if best is None else min(best, secs)
        cpu = {"value": sample * flops_per_transform(kind, n) / best / 1e9, "unit": "GFLOP/s", "cores": cores,
               "kind": impl,
               "sample": f"{sample} of the {batch} transforms (same synthetic inputs), batch split over {cores} host "
                         f"threads, one {'FFT' if kind == 'c2c' else 'RealFFT'}<{dtype}> object per thread, plan "
                         "build excluded, best of 2"}

    # ---- the other BASELINE configs, short runs (the headline fields above stay config 2)
    configs = None
    if args.workload == "c2" and not args.no_configs and not args.batch:
        del x, y
        torch.cuda.empty_cache()
        configs = {}
        t_cfg = time.perf_counter()
        names = ["c3", "c4-1000", "c4-2187", "c4-3125", "c4-6000", "c4-1000-f64", "c4-2187-f64", "c4-3125-f64", "c4-6000-f64"]
        for name in names:
            try:
                configs[name] = measure_batched(name, dev, rank, world, dist)
            except Exception as exc:  # a config that cannot run must not lose the headline
                configs[name] = {"error": repr(exc)[:300]}
            torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            try:
                configs["c1"] = measure_c1(dev)
            except Exception as exc:
                configs["c1"] = {"error": repr(exc)[:300]}
        if world > 1:
            try:
                c5 = measure_dist(args, 1 << 30, rank, world, dev, dist, steps=3, warmup=2, verify=True)
                if rank == 0:
                    configs["c5"] = {"workload": c5["config"]["workload"], "ms": c5["ms_per_step"], "gflops": c5["value"],
                                     "phase_ms_rank0": c5["config"]["phase_ms_rank0"],
                                     "nvlink_GBps_per_gpu_if_exchange_only": c5["roofline"]["nvlink_GBps_per_gpu_if_exchange_only"],
                                     "checks": c5.get("checks")}
            except Exception as exc:
                configs["c5"] = {"error": repr(exc)[:300]}
            # the same transform through the C ABI of the single-process plan (ssfft_dist_*): rank 0 drives all the
            # GPUs of the node over peer access while the other ranks wait
            try:
                torch.cuda.empty_cache()
                torch.cuda.synchronize()
                # the other ranks wait on the CPU (gloo): an NCCL barrier would keep a spinning kernel on their GPUs
                # while rank 0 is timing a transform that uses those GPUs
                cpu_group = dist.new_group(backend="gloo")
                dist.barrier(group=cpu_group)
                if rank == 0:
                    configs["c5_c_abi"] = measure_dist_local(1 << 30, world)
                dist.barrier(group=cpu_group)
            except Exception as exc:
                configs["c5_c_abi"] = {"error": repr(exc)[:300]}
        configs["seconds"] = round(time.perf_counter() - t_cfg, 1)

    traffic, traffic_src = profiled_traffic(args.workload) if not args.batch else (None, None)
    if rank == 0:
        line = {
            "metric": metric_for(args.workload), "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if dtype == "float32" else "f64",
            "data": f"synthetic uniform[-0.5,0.5), counter-based generator (device twin of the oracle's), seed {SEED}",
            "config": {"workload": f"{args.workload}: {kind} {dtype} N={n} x {batch} transforms per GPU"
                                   + (" (forward + inverse per step)" if kind != "c2c" else " (forward)"),
                       "plan": plan.describe(), "parallelism": f"batch-sharded x{world}, no collective",
                       "l2": f"working set {2 * batch * n * (2 if kind == 'c2c' else 1) * esz / 2**20:.0f} MiB per GPU "
                             ">> 126 MB L2 (inputs larger than L2, no flush needed)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_step_per_gpu": bytes_step_gpu,
                         "kernel_ms": ms_step},
            "clocks": clocks, "gpu_launches": launches,
        }
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        if configs:
            line["configs"] = configs
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
