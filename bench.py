#!/usr/bin/env python
"""Benchmark of the SSL hot path: two-view augmentation of 16-bit slices + NT-Xent fwd/bwd.

    python bench.py --gpus N --steps K --warmup W                   # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path (oracle port)
    python bench.py --config cfg2|cfg3|cfg4|cfg5 ...                # BASELINE.json configs[1..4]

Default workload = the configuration BASELINE.json's metric is quoted on: GLOBAL batch 4096 (cfg3), strong scaling:
4096 synthetic 512x512 uint16 slices over the N ranks -> 8192 views 224x224 bf16 (kernel K1), then NT-Xent (T = 0.1)
forward + backward over the 8192 x 128 embeddings, all-gathered across ranks when N > 1 (kernels K2/K3).  The
backbone between the views and the embeddings is out of scope (stock PyTorch/cuDNN in the reference), so the
embeddings are a fixed synthetic [2B, D] tensor with requires_grad.

One "step" = one pass of the hot path over one batch.  Prints ONE JSON line (rank 0).  `value` = views/s of the whole
job with the slices resident in HBM; `e2e` = the same through the public API from pinned HOST slices (H2D inside
the timed region, loss read back); `roofline` = kernel K1 (the dominant kernel) against measured HBM bandwidth;
`cpu_baseline` = the oracle's torchvision chain + CPU NT-Xent on this box's host cores.  At N > 1 the NVLink exchange
is first checked against a float64 torch restatement of the rank-sharded loss (`exchange_parity`).
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
METRIC = "aug views/sec + NT-Xent fwd+bwd ms at global batch 4096"

# BASELINE.json configs (1-based like SURVEY 8d): global batch (0 = per-GPU batch, weak), crop, D
CONFIGS = {
    "cfg2": dict(global_batch=0, batch=1024, crop=224, dim=128, scaling="weak",
                 name="cfg2 (1xB200: fused aug + NT-Xent, batch 1024 per GPU, 224x224 crops, 128-d proj)"),
    "cfg3": dict(global_batch=4096, batch=0, crop=224, dim=128, scaling="strong",
                 name="cfg3 (SimCLR-style step, global batch 4096 with cross-GPU embedding all-gather)"),
    "cfg4": dict(global_batch=16384, batch=0, crop=256, dim=2048, scaling="strong",
                 name="cfg4 (large-batch stress, global batch 16384, 2048-d proj, 256x256 crops)"),
    "cfg5": dict(global_batch=0, batch=4096, crop=224, dim=128, scaling="weak",
                 name="cfg5 (fused aug only, 512x512 CT windowing/normalise at 96x96 and 224x224 crops, HBM GB/s sweep)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override: images per GPU (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0, help="override: total images over all GPUs (strong scaling)")
    ap.add_argument("--image", type=int, default=512)
    ap.add_argument("--crop", type=int, default=0)
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--temperature", type=float, default=0.1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling (1024 slices/GPU) measurement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg.update(batch=args.batch, global_batch=0, scaling="weak")
    if args.global_batch:
        cfg.update(global_batch=args.global_batch, batch=0, scaling="strong")
    if args.crop:
        cfg["crop"] = args.crop
    if args.dim:
        cfg["dim"] = args.dim
    args.cfg = cfg
    return args


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
    cfg = args.cfg
    crop, dim = cfg["crop"], cfg["dim"]
    batch = 256                     # BASELINE.json configs[0]: a bounded sample of the workload the CPU can finish
    steps = max(1, min(args.steps, 12))
    warm = max(1, min(args.warmup, 2))
    per_step_budget = max(1.0, min(8.0, 150.0 / max(1, steps + warm)))
    r = cpu_reference_step_rate(args.image, crop, dim, args.temperature, batch,
                                budget_s=per_step_budget * steps, warmup_steps=warm, max_steps=steps)
    sample = (f"{r['batches']} steps of {batch} slices {args.image}x{args.image} -> {2 * batch} views {crop}^2 "
              f"(torchvision v2 chain, one worker process per core, {r['cores']} cores) + NT-Xent fwd+bwd fp32 on "
              f"{2 * batch}x{dim}; a bounded sample of the {cfg['name']} workload (views/s does not depend on the batch)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["views_per_s"], "unit": "views/s", "n_gpus": args.gpus,
        "steps": r["batches"], "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform uint16",
        "config": {"workload": f"CPU reference on {cfg['name']}: two-view aug + NT-Xent (T={args.temperature}), "
                               f"sample batch {batch}, {args.image}x{args.image} uint16 -> {crop}^2, D={dim}"},
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
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        # nvidia-smi attaches to every GPU of the box while it starts (~1 s on an 8-GPU node) and that slows kernel launches
        # of all ranks; wait for its first sample so the timed region sees only the steady 50 ms polling
        self.first = ""
        try:
            import select
            if select.select([self.proc.stdout], [], [], 15.0)[0]:
                self.first = self.proc.stdout.readline()
        except Exception:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        out = getattr(self, "first", "") + (out or "")
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def load_traffic(batch, crop):
    """dram bytes per K1 launch from the committed `ncu --set full` capture of the SAME launch (slices per GPU, crop,
    current kernel), else None: the figure cannot be measured inside an un-profiled run."""
    path = os.path.join(ROOT, "profiles", "aug_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            t = json.load(f)
        key = f"B{batch}_s{crop}"
        if key in t and t[key].get("kernel") == "aug_strip_kernel":
            return t[key]["dram_bytes_per_launch"]
    return None


def exchange_parity(dist, torch, dev, rank, world, shapes, temperature):
    """Multi-rank parity of the loss (NVLink exchange + option-L backward) against a float64 torch restatement of the
    rank-sharded convention: rank r's loss is the mean over its rows against all columns, its gradient
    sum_r' dL_r'/dz_local.  Returns (ok on every rank, worst loss rel err, worst grad Frobenius rel err)."""
    import torch.nn.functional as F

    from medical_image_segmentation_b200 import nt_xent_rows
    ok, worst_l, worst_g = True, 0.0, 0.0
    for (b_local, D) in shapes:
        rows = 2 * b_local
        g = torch.Generator(device=dev).manual_seed(4242 + rank)
        z = torch.randn(rows, D, device=dev, generator=g)
        z_all = torch.empty(world * rows, D, device=dev)
        dist.all_gather_into_tensor(z_all, z)
        z64 = z_all.double().requires_grad_(True)
        u = F.normalize(z64, dim=1)
        s = (u @ u.T) / temperature
        n2 = world * rows
        s = s.masked_fill(torch.eye(n2, dtype=torch.bool, device=dev), float("-inf"))
        idx = torch.arange(n2, device=dev)
        pos = (idx // rows) * rows + ((idx % rows) + rows // 2) % rows
        per_row = F.cross_entropy(s, pos, reduction="none")
        losses = per_row.view(world, rows).mean(dim=1)
        losses.sum().backward()
        ref_loss = float(losses[rank])
        ref_grad = z64.grad[rank * rows:(rank + 1) * rows]
        zz = z.clone().requires_grad_(True)
        loss = nt_xent_rows(zz, temperature, dist.group.WORLD)
        loss.backward()
        lrel = abs(float(loss) - ref_loss) / abs(ref_loss)
        fro = float((zz.grad.double() - ref_grad).norm() / ref_grad.norm())
        ok = ok and lrel <= 1e-3 and fro <= 1e-3
        worst_l, worst_g = max(worst_l, lrel), max(worst_g, fro)
        del z64, u, s, per_row
    t = torch.tensor([1.0 if ok else 0.0, -worst_l, -worst_g], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t[0].item() == 1.0), -float(t[1]), -float(t[2])


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MIS_NTXENT_GRAPH", "1")      # the loss's launches are replayed as CUDA graphs
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from medical_image_segmentation_b200 import FusedTwoViewTransforms, algorithmic_bytes, nt_xent_rows, peer
    from medical_image_segmentation_b200.loss import CudaKernels

    cfg = args.cfg
    B = cfg["batch"] if not cfg["global_batch"] else cfg["global_batch"] // world
    H = W = args.image
    s, D = cfg["crop"], cfg["dim"]
    group = dist.group.WORLD if world > 1 else None
    hbm_peak, tc_peak, peak_src = load_peaks()

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

    # ---- multi-rank parity of the exchange, before anything is timed ---------------------------------------------
    parity = None
    if world > 1:
        shapes = [(64, 64), (B, D)] if args.config != "cfg5" else [(64, 64)]
        ok, lrel, fro = exchange_parity(dist, torch, dev, rank, world, shapes, args.temperature)
        parity = {"status": "PASS" if ok else "FAIL", "loss_rel": lrel, "grad_fro_rel": fro,
                  "shapes": [f"{2 * b}x{d} per rank" for b, d in shapes],
                  "reference": "float64 torch restatement of the rank-sharded NT-Xent (SURVEY A.5), all ranks",
                  "transport": "peer" if peer._cache else "nccl"}
        if not ok:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "exchange_parity": parity, "error": "multi-rank parity failed"}), flush=True)
            dist.destroy_process_group()
            raise SystemExit(3)

    # ---- synthetic inputs (device resident for `value`, pinned host copy for `e2e`) ---------------
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = torch.randint(0, 65536, (B, 1, H, W), dtype=torch.int32, device=dev, generator=g).to(torch.uint16)
    z = torch.randn(2 * B, D, device=dev, generator=g).requires_grad_(True)
    t = FusedTwoViewTransforms(s, (MEAN,), (STD,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), prefetch_params=True,
                               generator=torch.Generator().manual_seed(1000 + rank))   # host RNG replay of step k+1 overlaps step k
    out = torch.empty((2 * B, 1, s, s), dtype=torch.bfloat16, device=dev)
    torch.manual_seed(1000 + rank)
    aug_only = args.config == "cfg5"

    def step_device():
        params = t.next_params(B, H, W, view_major=True)       # host RNG replay, same stream as the reference
        t.apply(x_dev, params, out)
        if aug_only:
            return None
        z.grad = None
        loss = nt_xent_rows(z, args.temperature, group)
        loss.backward()
        return loss

    # ---- the clock sampler runs from here to the end of the timed steps: its start-up (nvidia-smi attaches to every GPU
    #      of the box, ~1 s) must not fall into the timed region, and the process must not sit idle right before it --------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                  # polls clocks / throttle reasons every 50 ms from here on
    if world > 1:
        dist.barrier()

    # ---- per-kernel breakdown first (it also brings clocks, allocator and exchange buffers to steady state) ---------
    bsteps = max(5, min(args.steps, 50))
    torch.manual_seed(7)
    params = t.to_view_major(t.draw_params(B, H, W))
    alg_bytes = algorithmic_bytes(params, 1, s)
    for _ in range(3):
        t.apply(x_dev, params, out)
    ms_aug = timed(lambda: t.apply(x_dev, params, out), bsteps) / bsteps
    achieved = alg_bytes / (ms_aug * 1e-3) / 1e9

    def loss_only():
        z.grad = None
        nt_xent_rows(z, args.temperature, group).backward()

    ms_loss = None
    if not aug_only:
        for _ in range(3):
            loss_only()
        ms_loss = timed(loss_only, bsteps) / bsteps
    n2 = 2 * B * world
    flops_rank = 6.0 * n2 * n2 * D / world

    sweep = None
    if aug_only and rank == 0:
        sweep = []
        for crop in (96, 224):
            for window in (None, (1000.0, 30000.0)):
                tt = FusedTwoViewTransforms(crop, (MEAN,), (STD,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), window=window)
                torch.manual_seed(7)
                pp = tt.to_view_major(tt.draw_params(B, H, W))
                oo = torch.empty((2 * B, 1, crop, crop), dtype=torch.bfloat16, device=dev)
                for _ in range(3):
                    tt.apply(x_dev, pp, oo)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(bsteps):
                    tt.apply(x_dev, pp, oo)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / bsteps
                ab = algorithmic_bytes(pp, 1, crop)
                sweep.append({"crop": crop, "window": list(window) if window else None, "slices": B, "ms": ms,
                              "views_per_s": 2 * B / (ms * 1e-3), "gbs": ab / (ms * 1e-3) / 1e9,
                              "frac": ab / (ms * 1e-3) / 1e9 / hbm_peak})
                del oo

    # ---- the timed region: W warm-up steps, then exactly K steps ---------------------------------------------------
    torch.manual_seed(1000 + rank)
    t.drain_prefetch()
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        step_device()
    launches0 = CudaKernels.launches + t.launches
    ms_total = timed(step_device, args.steps)
    gpu_launches = CudaKernels.launches + t.launches - launches0      # kernels of libmis_b200.so enqueued in the timed region
    ms_step = ms_total / args.steps
    value = world * 2 * B / (ms_step * 1e-3)
    # a longer look at the same step (>= 1 s) while the clock sampler is still running
    long_steps = int(max(args.steps, min(5000, 1000.0 / max(ms_step, 1e-3))))
    ms_long = timed(step_device, long_steps) / long_steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- weak-scaling companion (1024 slices per GPU, cfg2 per GPU) when the main run is strong scaling --------------
    weak = None
    if cfg["scaling"] == "strong" and B != 1024 and not args.no_weak and not aug_only and D <= 256:
        Bw = 1024
        xw = x_dev[:Bw] if B >= Bw else torch.randint(0, 65536, (Bw, 1, H, W), dtype=torch.int32, device=dev).to(torch.uint16)
        zw = torch.randn(2 * Bw, D, device=dev).requires_grad_(True)
        ow = out[:2 * Bw] if B >= Bw else torch.empty((2 * Bw, 1, s, s), dtype=torch.bfloat16, device=dev)

        def step_weak():
            p = t.next_params(Bw, H, W, view_major=True)
            t.apply(xw, p, ow)
            zw.grad = None
            nt_xent_rows(zw, args.temperature, group).backward()

        for _ in range(max(n_warm, 5)):
            step_weak()
        wsteps = max(20, min(args.steps, 200))
        ms_w = timed(step_weak, wsteps) / wsteps
        weak = {"images_per_gpu": Bw, "global_batch": Bw * world, "steps": wsteps, "ms_per_step": ms_w,
                "value": world * 2 * Bw / (ms_w * 1e-3), "unit": "views/s", "scaling": "weak"}
        t.drain_prefetch()

    # ---- e2e: pinned host slices -> H2D -> K1 -> NT-Xent fwd+bwd -> loss read back ---------------------
    e2e = None
    if not args.no_e2e:
        from medical_image_segmentation_b200.numa import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local)              # pinned staging memory on the GPU's own socket (first touch)
        x_host = torch.empty((B, 1, H, W), dtype=torch.uint16).pin_memory()
        x_host.copy_(x_dev)
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        h2d_bytes = []

        # the batch of step k+1 crosses PCIe on a copy stream, into the other of two device buffers, while step k computes
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [torch.empty_like(x_dev) for _ in range(2)]
        free_ev = [torch.cuda.Event(), torch.cuda.Event()]     # K1 has finished reading buffer i
        state = {"k": 0, "next": None}

        def issue(k):
            rec = t.next_params(B, H, W)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free_ev[k % 2])
                xs = t.stage_needed_rows(x_host, rec, dev, out=bufs[k % 2])   # only the rows the crops read (public API path)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return rec, xs, ev, t.last_h2d_bytes

        def step_e2e():
            if state["next"] is None:
                state["next"] = issue(state["k"])
            rec, xs, ev, nb = state["next"]
            k = state["k"]
            state["next"] = issue(k + 1)                   # enqueued before this step's kernels: the two overlap
            state["k"] = k + 1
            h2d_bytes.append(nb)
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            t.apply(xs, t.to_view_major(rec), out)
            free_ev[k % 2].record(cur)
            if aug_only:
                loss_host.copy_(out[0, 0, 0, 0].float(), non_blocking=True)
            else:
                z.grad = None
                loss = nt_xent_rows(z, args.temperature, group)
                loss.backward()
                loss_host.copy_(loss.detach(), non_blocking=True)
            cur.synchronize()                               # the caller consumes the loss every step
            return float(loss_host)

        for _ in range(3):
            step_e2e()
        e2e_steps = max(3, min(args.steps, 10))
        ms_e2e = timed(step_e2e, e2e_steps) / e2e_steps
        e2e = {"value": world * 2 * B / (ms_e2e * 1e-3), "unit": "views/s", "ms_per_step": ms_e2e, "steps": e2e_steps,
               "h2d_bytes_per_step": int(sum(h2d_bytes[-e2e_steps:]) / e2e_steps + params.nbytes), "d2h_bytes_per_step": 4,
               "h2d_gbs_per_gpu": sum(h2d_bytes[-e2e_steps:]) / e2e_steps / (ms_e2e * 1e-3) / 1e9, "numa": numa,
               "pipeline": "two device buffers: the H2D copy of step k+1 runs on a copy stream under the kernels of step k; "
                           "every step still issues one batch copy and reads its loss back",
               "h2d_note": f"rows no crop reads are skipped when the gap exceeds 256 KB "
                           f"({sum(h2d_bytes[-e2e_steps:]) / e2e_steps / (x_host.numel() * 2):.0%} of the {x_host.numel() * 2} B batch moved)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_step_rate(H, s, D if D <= 256 else 128, args.temperature, 256, budget_s=args.cpu_seconds)
        cpu = {"value": r["views_per_s"], "unit": "views/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['batches']} batches of 256 slices ({r['seconds']:.1f} s): torchvision v2 chain in "
                         f"{r['cores']} worker processes + NT-Xent fwd+bwd fp32 CPU ({r['ntxent_ms']:.1f} ms)",
               "ntxent_fwd_bwd_ms": r["ntxent_ms"]}

    peer.check_health()                                  # no consumer ever gave up waiting for a peer's flag
    if rank == 0:
        x_bytes = B * H * W * 2
        line = {
            "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic uniform uint16 slices, randn embeddings",
            "config": {"workload": f"{cfg['name']}: {B} slices/GPU {H}x{W} u16 -> {2 * B} views {s}x{s} bf16"
                                   + ("" if aug_only else f", NT-Xent over {n2}x{D}, T={args.temperature}"),
                       "name": args.config, "global_batch": B * world, "images_per_gpu": B, "crop": s, "proj_dim": D,
                       "l2": f"inputs ({x_bytes / 2 ** 20:.0f} MiB of slices per GPU, every step reads new random crops) "
                             + ("larger than the 126 MB L2; no explicit flush" if x_bytes > 2 * 126e6 else
                                "NOT much larger than L2: treat with care"),
                       "ntxent_operands": "tf32 (tcgen05 kind::tf32), fp32 accumulate",
                       "ntxent_launch": CudaKernels.launch_mode(world),
                       "exchange": ("none (single rank)" if world == 1 else
                                    ("NVLink peer stores fused into the producing kernels" if peer._cache else "NCCL all-gather"))},
            "sustained": {"steps": long_steps, "ms_per_step": ms_long, "value": world * 2 * B / (ms_long * 1e-3)},
            "aug_ms": ms_aug, "aug_views_per_s_per_gpu": 2 * B / (ms_aug * 1e-3),
            "ntxent_fwd_bwd_ms": ms_loss,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": load_traffic(B, s), "kernel": "aug_strip_kernel (K1)",
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "timing": f"CUDA events around {bsteps} back-to-back launches on the launching stream"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
        }
        if ms_loss is not None:
            # kind::tf32 runs at half the bf16 rate: the tensor peak for these kernels is half the measured bf16 figure
            tf32_peak = tc_peak / 2.0
            line["roofline_ntxent"] = {
                "bound": "tensor", "achieved": flops_rank / (ms_loss * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": flops_rank / (ms_loss * 1e-3) / 1e12 / tf32_peak,
                "note": "6*(2N)^2*D/world algorithmic flops over fwd+bwd time incl. launch overhead and (N>1) both "
                        "exchanges; peak = tf32 dense = half the measured bf16 cuBLAS figure (kernels run kind::tf32)"}
        if parity is not None:
            line["exchange_parity"] = parity["status"]
            line["exchange_parity_detail"] = parity
        if weak is not None:
            line["weak_scaling"] = weak
        if sweep is not None:
            line["aug_sweep"] = sweep
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
