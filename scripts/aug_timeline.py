"""Per-phase clock timeline of K1 CTAs (uses the mis_debug_set_stamp_buffer profiling aid)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from medical_image_segmentation_b200 import _lib
from medical_image_segmentation_b200.transforms import FusedTwoViewTransforms
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 224
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randint(0, 65536, (B, 1, 512, 512), dtype=torch.int32, device="cuda", generator=g).to(torch.uint16)
import os
t = FusedTwoViewTransforms(crop, (0.227358,), (0.237160,), use_tma=int(os.environ.get("TMA", "0")))
torch.manual_seed(0)
params = t.to_view_major(t.draw_params(B, 512, 512))
out = torch.empty((2 * B, 1, crop, crop), dtype=torch.bfloat16, device="cuda")
nb = (crop + 31) // 32
grid = nb * 2 * B
for _ in range(2):
    t.apply(x, params, out)
buf = torch.zeros((grid, 8), dtype=torch.int64, device="cuda")
_lib.lib.mis_debug_set_stamp_buffer(buf.data_ptr())
t.apply(x, params, out)
torch.cuda.synchronize()
_lib.lib.mis_debug_set_stamp_buffer(None)
s = buf.cpu().numpy().astype(np.float64)
d = np.diff(s[:, :7], axis=1)
names = ["tables+sched", "V pass (warp0)", "wait other V warps", "H pass", "colour+cluster", "store"]
tot = s[:, 6] - s[:, 0]
print(f"CTAs {grid}; CTA duration clk: mean {tot.mean():.0f} median {np.median(tot):.0f} p90 {np.percentile(tot,90):.0f}")
for i, n in enumerate(names):
    print(f"  {n:22s} mean {d[:, i].mean():8.0f} clk  ({100*d[:, i].mean()/tot.mean():5.1f}%)  median {np.median(d[:, i]):8.0f}")
first = s[:, 7] - s[:, 1]
print(f"  [V pass] first row landed after {first.mean():.0f} clk (median {np.median(first):.0f}) from V-pass start")
span = s[:, 6].max() - s[:, 0].min()
print("kernel span (clk, across SMs; clocks not synchronised between SMs):", span)
