#!/usr/bin/env python
"""Benchmark of the SSL hot path: two-view augmentation of 16-bit slices + NT-Xent fwd/bwd.

    python bench.py --gpus N --steps K --warmup W                 # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path (oracle port)

One "step" = one pass of the hot path over one batch (BASELINE.json configs[1] per GPU):
    B = 1024 synthetic 512x512 uint16 slices  ->  2048 views 224x224 bf16 (kernel K1)
    NT-Xent (T = 0.1) forward + backward over the 2048 x 128 embeddings of the batch (kernels K2/K3),
    all-gathered across ranks when N > 1 (weak scaling: per-GPU batch fixed, global batch = 1024*N).
The backbone between the views and the embeddings is out of scope (stock PyTorch/cuDNN in the
reference), so the embeddings are a fixed synthetic [2B, D] tensor with requires_grad.

Prints ONE JSON line (rank 0).  `value` = views/s of the whole job with the slices resident in HBM;
`e2e` = the same through the public API from pinned HOST slices (H2D inside the timed region, loss
read back); `roofline` = kernel K1 (the dominant kernel) against measured HBM bandwidth;
`cpu_baseline` = the oracle's torchvision chain + CPU NT-Xent on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MEAN, STD = 57.9764 / 255.0, 60.4759 / 255.0      # lightning_module.py:212-213 on the [0,1] scale
METRIC = "aug views/sec + NT-Xent fwd+bwd ms"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU")
    ap.add_argument("--global-batch", type=int, default=0, help="strong-scaling variant: total images over all GPUs")
    ap.add_argument("--image", type=int, default=512)
    ap.add_argument("--crop", type=int, default=224)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--temperature", type=float, default=0.1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's path (oracle port: torchvision chain per sample in worker processes)
# ------------------------------------------------------------------------------------------------
_W = {}


def _worker_init(image, crop):
    import torch
    from oracle.aug_oracle import TwoViewChainTV
    torch.set_num_threads(1)                       # the reference parallelises across worker processes
    g = torch.Generator().manual_seed(1234)
    _W["x"] = torch.randint(0, 65536, (8, image, image), generator=g, dtype=torch.int32)
    _W["chain"] = TwoViewChainTV(crop, (MEAN,), (STD,), (0.0, 0.0), (0.0, 0.0))


def _worker_chunk(job):
    """Augment `n` slices exactly as a DataLoader worker of the reference does: one sample at a time."""
    import torch
    from torchvision import tv_tensors
    seed, n = job
    torch.manual_seed(seed)
    acc = 0.0
    for i in range(n):
        x = tv_tensors.Image((_W["x"][i % 8].to(torch.float32) * (1.0 / 65535.0))[None])
        v1, v2 = _W["chain"](x)
        acc += float(v1[0, 0, 0]) + float(v2[0, 0, 0])
    return acc


def cpu_reference_step_rate(image, crop, dim, temperature, batch, budget_s, warmup_steps=1, max_steps=10 ** 9):
    """views/s of the reference path on the host.  A step = the two views of `batch` slices produced by one worker
    process per core (torchvision v2 chain, per sample, 1 thread each) followed by NT-Xent fwd+bwd (fp32 torch, all
    cores) on the batch's embeddings -- the two legs are serial, so neither competes with the other for cores."""
    import multiprocessing as mp

    import torch
    from oracle.loss_oracle import ntxent_loss
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    z1 = torch.randn(batch, dim, generator=g)
    z2 = torch.randn(batch, dim, generator=g)
    nw = min(cores, batch)
    jobs = [(1000 + k, batch // nw + (1 if k < batch % nw else 0)) for k in range(nw)]
    step_ms, loss_ms = [], []
    with mp.get_context("spawn").Pool(nw, initializer=_worker_init, initargs=(image, crop)) as pool:
        t_begin = None
        k = 0
        while True:
            t0 = time.perf_counter()
            pool.map(_worker_chunk, jobs, chunksize=1)
            a = z1.clone().requires_grad_(True)
            b = z2.clone().requires_grad_(True)
            t1 = time.perf_counter()
            ntxent_loss(a, b, temperature).backward()
            t2 = time.perf_counter()
            if k >= warmup_steps:
                step_ms.append((t2 - t0) * 1e3)
                loss_ms.append((t2 - t1) * 1e3)
            elif k == warmup_steps - 1:
                t_begin = time.perf_counter()
            k += 1
            if len(step_ms) >= max_steps or (t_begin is not None and time.perf_counter() - t_begin >= budget_s
                                             and len(step_ms) >= 1):
                break
    ms = statistics.mean(step_ms)
    return dict(views_per_s=2 * batch / (ms * 1e-3), seconds=sum(step_ms) * 1e-3, batches=len(step_ms), cores=nw,
                ntxent_ms=statistics.median(loss_ms), ms_per_step=ms)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 256                                       # BASELINE.json configs[0]: the reference's CPU-runnable case
    per_step_budget = max(1.0, min(8.0, 150.0 / max(1, args.steps + args.warmup)))
    r = cpu_reference_step_rate(args.image, args.crop, args.dim, args.temperature, batch,
                                budget_s=per_step_budget * args.steps, warmup_steps=max(1, args.warmup),
                                max_steps=args.steps)
    sample = (f"{r['batches']} steps of {batch} slices {args.image}x{args.image} -> {2 * batch} views {args.crop}^2 "
              f"(torchvision v2 chain, one worker process per core, {r['cores']} cores) + NT-Xent fwd+bwd fp32 on {2 * batch}x{args.dim}")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["views_per_s"], "unit": "views/s", "n_gpus": args.gpus,
        "steps": r["batches"], "warmup": max(1, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform uint16",
        "config": {"workload": f"CPU reference: two-view aug + NT-Xent (T={args.temperature}), batch {batch}, "
                               f"{args.image}x{args.image} uint16 -> {args.crop}^2, D={args.dim}"},
        "ntxent_fwd_bwd_ms": r["ntxent_ms"],
        "cpu_baseline": {"value": r["views_per_s"], "unit": "views/s", "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["views_per_s"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.idx = None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def load_traffic(batch, crop):
    """dram bytes per K1 launch from the committed ncu capture of the same workload, else None."""
    path = os.path.join(ROOT, "profiles", "aug_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            t = json.load(f)
        key = f"B{batch}_s{crop}"
        if key in t:
            return t[key]["dram_bytes_per_launch"]
    return None


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MIS_NTXENT_GRAPH", "1")      # single-rank loss: the seven launches replayed as one CUDA graph
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from medical_image_segmentation_b200 import FusedTwoViewTransforms, algorithmic_bytes, nt_xent_rows, peer
    from medical_image_segmentation_b200.loss import CudaKernels

    B = args.batch if not args.global_batch else args.global_batch // world
    H = W = args.image
    s, D = args.crop, args.dim
    group = dist.group.WORLD if world > 1 else None

    # ---- synthetic inputs (device resident for `value`, pinned host copy for `e2e`) ---------------
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = torch.randint(0, 65536, (B, 1, H, W), dtype=torch.int32, device=dev, generator=g).to(torch.uint16)
    z = torch.randn(2 * B, D, device=dev, generator=g).requires_grad_(True)
    t = FusedTwoViewTransforms(s, (MEAN,), (STD,), prefetch_params=True)   # host RNG replay of step k+1 overlaps step k
    out = torch.empty((2 * B, 1, s, s), dtype=torch.bfloat16, device=dev)
    torch.manual_seed(1000 + rank)

    def step_device():
        params = t.to_view_major(t.next_params(B, H, W))      # host RNG replay, same stream as the reference
        t.apply(x_dev, params, out)
        z.grad = None
        loss = nt_xent_rows(z, args.temperature, group)
        loss.backward()
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return ms

    # warm-up: the requested W (>= 3) steps, topped up to 50 -- the first ~40 steps of a fresh process run ~8 % slow
    # (clock ramp, allocator and peer-exchange set-up); the timed region is exactly args.steps steps after that
    n_warm = max(args.warmup, 3, 50)
    for _ in range(n_warm):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = CudaKernels.launches + t.launches
    ms_total = timed(step_device, args.steps)
    gpu_launches = CudaKernels.launches + t.launches - launches0      # kernels of libmis_b200.so enqueued in the timed region
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * 2 * B / (ms_step * 1e-3)

    # ---- per-kernel breakdown (same stream, CUDA events, inputs 512 MiB > L2) ------------------------
    t.drain_prefetch()
    torch.manual_seed(7)
    params = t.to_view_major(t.draw_params(B, H, W))
    alg_bytes = algorithmic_bytes(params, 1, s)
    for _ in range(3):
        t.apply(x_dev, params, out)
    ms_aug = timed(lambda: t.apply(x_dev, params, out), args.steps) / args.steps

    def loss_only():
        z.grad = None
        nt_xent_rows(z, args.temperature, group).backward()

    ms_loss = timed(loss_only, args.steps) / args.steps
    hbm_peak, tc_peak, peak_src = load_peaks()
    achieved = alg_bytes / (ms_aug * 1e-3) / 1e9
    n2 = 2 * B * world
    flops_rank = 6.0 * n2 * n2 * D / world

    # ---- e2e: pinned host slices -> H2D -> K1 -> NT-Xent fwd+bwd -> loss read back ---------------------
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty((B, 1, H, W), dtype=torch.uint16).pin_memory()
        x_host.copy_(x_dev)
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()

        h2d_bytes = []

        def step_e2e():
            rec = t.next_params(B, H, W)
            xs = t.stage_needed_rows(x_host, rec, dev)   # only the rows the crops read; near ranges merged (public API path)
            h2d_bytes.append(t.last_h2d_bytes)
            t.apply(xs, t.to_view_major(rec), out)
            z.grad = None
            loss = nt_xent_rows(z, args.temperature, group)
            loss.backward()
            loss_host.copy_(loss.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()       # the caller consumes the loss every step
            return float(loss_host)

        for _ in range(3):
            step_e2e()
        e2e_steps = max(3, min(args.steps, 10))
        ms_e2e = timed(step_e2e, e2e_steps) / e2e_steps
        e2e = {"value": world * 2 * B / (ms_e2e * 1e-3), "unit": "views/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(sum(h2d_bytes[-e2e_steps:]) / e2e_steps + params.nbytes), "d2h_bytes_per_step": 4,
               "h2d_note": f"rows no crop reads are skipped when the gap exceeds 256 KB "
                           f"({sum(h2d_bytes[-e2e_steps:]) / e2e_steps / (x_host.numel() * 2):.0%} of the {x_host.numel() * 2} B batch moved)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_step_rate(H, s, D, args.temperature, 256, budget_s=args.cpu_seconds)
        cpu = {"value": r["views_per_s"], "unit": "views/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['batches']} batches of 256 slices ({r['seconds']:.1f} s): torchvision v2 chain in "
                         f"{r['cores']} worker processes + NT-Xent fwd+bwd fp32 CPU ({r['ntxent_ms']:.1f} ms)",
               "ntxent_fwd_bwd_ms": r["ntxent_ms"]}

    peer.check_timeouts()                                # no consumer ever gave up waiting for a peer's flag
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic uniform uint16 slices, randn embeddings",
            "config": {"workload": f"{'8xB200 cfg3-style' if args.global_batch else '1xB200 cfg2 per GPU'}: fused aug + "
                                   f"NT-Xent, {B} slices/GPU {H}x{W} u16 -> {2 * B} views {s}x{s} bf16, D={D}, T={args.temperature}",
                       "global_batch": B * world, "images_per_gpu": B, "crop": s, "proj_dim": D,
                       "l2": "inputs (512 MiB/GPU) larger than L2; no explicit flush",
                       "ntxent_operands": "tf32 (tcgen05 kind::tf32), fp32 accumulate",
                       "ntxent_launch": ("CUDA graph replay (prep, fwd, bwd: 6 kernels + memset)"
                                         if world == 1 and os.environ.get("MIS_NTXENT_GRAPH") == "1" else "eager"),
                       "exchange": ("none (single rank)" if world == 1 else
                                    ("NVLink peer stores fused into the producing kernels" if peer._cache else "NCCL all-gather"))},
            "aug_ms": ms_aug, "aug_views_per_s_per_gpu": 2 * B / (ms_aug * 1e-3),
            "ntxent_fwd_bwd_ms": ms_loss,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": load_traffic(B, s), "kernel": "aug_tile_kernel (K1)",
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
            "roofline_ntxent": {"bound": "tensor", "achieved": flops_rank / (ms_loss * 1e-3) / 1e12, "peak": tc_peak,
                                "unit": "TFLOP/s", "frac": flops_rank / (ms_loss * 1e-3) / 1e12 / tc_peak,
                                "note": "6*(2N)^2*D/world algorithmic flops over fwd+bwd wall time incl. host launch "
                                        "overhead and (N>1) the all-gathers; peak = measured bf16, kernels run tf32"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
