"""Host-side duration of each of the first steps after a device synchronize (is there a ramp?)."""
import os, sys, time
import torch
sys.path.insert(0, ".")
os.environ.setdefault("MIS_NTXENT_GRAPH", "1")
from medical_image_segmentation_b200 import FusedTwoViewTransforms, nt_xent_rows

B, H, W, s, D = int(os.environ.get("HB", "256")), 512, 512, 224, 128
x = torch.randint(0, 65536, (B, 1, H, W), dtype=torch.int32, device="cuda").to(torch.uint16)
z = torch.randn(2 * B, D, device="cuda").requires_grad_(True)
t = FusedTwoViewTransforms(s, (0.227,), (0.237,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), prefetch_params=True,
                           generator=torch.Generator().manual_seed(1))
out = torch.empty((2 * B, 1, s, s), dtype=torch.bfloat16, device="cuda")
def step():
    vm = t.next_params(B, H, W, view_major=True)
    t.apply(x, vm, out)
    z.grad = None
    nt_xent_rows(z, 0.1, None).backward()
for _ in range(300):
    step()
def step_parts():
    a = time.perf_counter(); vm = t.next_params(B, H, W, view_major=True)
    b = time.perf_counter(); t.apply(x, vm, out)
    c = time.perf_counter(); z.grad = None; loss = nt_xent_rows(z, 0.1, None)
    d = time.perf_counter(); loss.backward()
    e = time.perf_counter()
    return [1e6 * (b - a), 1e6 * (c - b), 1e6 * (d - c), 1e6 * (e - d)]
for trial in range(3):
    torch.cuda.synchronize()
    parts = [step_parts() for _ in range(4)]
    torch.cuda.synchronize()
    print(f"parts {trial}: first steps after a sync, us [params, apply, loss fwd, loss bwd]:", [[round(v) for v in p] for p in parts])
for trial in range(3):
    torch.cuda.synchronize()
    if trial == 2:
        time.sleep(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = [time.perf_counter()]
    e0.record()
    for _ in range(40):
        step()
        ts.append(time.perf_counter())
    e1.record()
    torch.cuda.synchronize()
    d = [1e3 * (b - a) for a, b in zip(ts, ts[1:])]
    print(f"trial {trial}{' (after 1 s idle)' if trial == 2 else ''}: events {e0.elapsed_time(e1)/40:.4f} ms/step; host per step:",
          " ".join(f"{v:.2f}" for v in d))
