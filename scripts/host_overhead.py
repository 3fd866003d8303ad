"""Where the host time of one bench step goes (no device syncs inside the timed loops)."""
import os, sys, time
import torch
sys.path.insert(0, ".")
os.environ.setdefault("MIS_NTXENT_GRAPH", "1")
import torch.distributed as dist
from medical_image_segmentation_b200 import FusedTwoViewTransforms, nt_xent_rows

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = dist.group.WORLD

B, H, W, s, D = int(os.environ.get("HB", "1024")), 512, 512, 224, 128
x = torch.randint(0, 65536, (B, 1, H, W), dtype=torch.int32, device="cuda").to(torch.uint16)
z = torch.randn(2 * B, D, device="cuda").requires_grad_(True)
t = FusedTwoViewTransforms(s, (0.227,), (0.237,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), prefetch_params=True)
out = torch.empty((2 * B, 1, s, s), dtype=torch.bfloat16, device="cuda")
torch.manual_seed(rank)
acc = {"next_params": 0.0, "view_major": 0.0, "apply": 0.0, "loss_fwd": 0.0, "loss_bwd": 0.0}
def step(timing):
    t0 = time.perf_counter(); vm = t.next_params(B, H, W, view_major=True)
    t1 = time.perf_counter()
    t2 = time.perf_counter(); t.apply(x, vm, out)
    t3 = time.perf_counter(); z.grad = None; loss = nt_xent_rows(z, 0.1, group)
    t4 = time.perf_counter(); loss.backward()
    t5 = time.perf_counter()
    if timing:
        for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            acc[k] += v
for _ in range(10):
    step(False)
torch.cuda.synchronize()
n = 200
w0 = time.perf_counter()
for _ in range(n):
    step(True)
w1 = time.perf_counter()
torch.cuda.synchronize()
w2 = time.perf_counter()
msg = f"rank {rank}/{world}: host enqueue {1e3*(w1-w0)/n:.3f} ms/step, with drain {1e3*(w2-w0)/n:.3f} ms/step | " + ", ".join(f"{k} {1e6*v/n:.0f}us" for k, v in acc.items())
print(msg, flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
