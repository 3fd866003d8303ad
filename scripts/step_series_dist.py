"""Per-bin step time of the first steps after a barrier, on every rank (what is slow at the start of a timed region?).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/step_series_dist.py
"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
os.environ.setdefault("MIS_NTXENT_GRAPH", "1")
from medical_image_segmentation_b200 import FusedTwoViewTransforms, nt_xent_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
group = dist.group.WORLD
B, H, W, s, D = 4096 // world, 512, 512, 224, 128
x = torch.randint(0, 65536, (B, 1, H, W), dtype=torch.int32, device="cuda").to(torch.uint16)
z = torch.randn(2 * B, D, device="cuda").requires_grad_(True)
out = torch.empty((2 * B, 1, s, s), dtype=torch.bfloat16, device="cuda")


def make(prefetch):
    return FusedTwoViewTransforms(s, (0.227,), (0.237,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0),
                                  prefetch_params=prefetch, generator=torch.Generator().manual_seed(1000 + rank))


def run(name, step, nbins=12, per=20, pre=None):
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    if pre:
        pre()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(nbins + 1)]
    ts = [time.perf_counter()]
    ev[0].record()
    for b in range(nbins):
        for _ in range(per):
            step()
        ev[b + 1].record()
        ts.append(time.perf_counter())
    torch.cuda.synchronize()
    gpu = torch.tensor([ev[b].elapsed_time(ev[b + 1]) / per for b in range(nbins)], device="cuda")
    host = torch.tensor([1e3 * (ts[b + 1] - ts[b]) / per for b in range(nbins)], device="cuda")
    gmax, hmax = gpu.clone(), host.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name}: ms/step per {per}-step bin, max over ranks")
        print("   device:", " ".join(f"{v:.3f}" for v in gmax.tolist()))
        print("   host  :", " ".join(f"{v:.3f}" for v in hmax.tolist()), flush=True)


t = make(True)
def full():
    p = t.next_params(B, H, W, view_major=True)
    t.apply(x, p, out)
    z.grad = None
    nt_xent_rows(z, 0.1, group).backward()
t2 = make(False)
def full_noprefetch():
    p = t2.to_view_major(t2.draw_params(B, H, W))
    t2.apply(x, p, out)
    z.grad = None
    nt_xent_rows(z, 0.1, group).backward()
fixed = t.to_view_major(t.draw_params(B, H, W))
def fixed_table():
    t.apply(x, fixed, out)
    z.grad = None
    nt_xent_rows(z, 0.1, group).backward()
def loss_only():
    z.grad = None
    nt_xent_rows(z, 0.1, group).backward()
def aug_only():
    p = t.next_params(B, H, W, view_major=True)
    t.apply(x, p, out)

for _ in range(100):
    full()
run("full step", full)
run("full step again", full)
run("full step after 1 s idle", full, pre=lambda: time.sleep(1.0))
run("no prefetch thread", full_noprefetch)
run("fixed table (no host RNG)", fixed_table)
run("loss only", loss_only)
run("aug only", aug_only)
run("full step, last", full)
dist.barrier()
dist.destroy_process_group()
