"""HBM-bound scan check for mis_u16_moments: GB/s against the measured copy bandwidth."""
import json, os, sys
import torch
sys.path.insert(0, ".")
from medical_image_segmentation_b200.metrics import accumulate_moments
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = torch.randint(0, 65536, (B, 1, 512, 512), dtype=torch.int32, device="cuda").to(torch.uint16)
sums = torch.zeros((1, 2), dtype=torch.int64, device="cuda")
for _ in range(3):
    accumulate_moments(x, sums)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    accumulate_moments(x, sums)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gbs = x.numel() * 2 / ms / 1e6
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
print(f"moments over {x.numel()*2/2**30:.2f} GiB: {ms:.3f} ms, {gbs:.0f} GB/s read = {100*gbs/peak:.1f}% of the measured copy bandwidth {peak} "
      f"(a copy counts read+write bytes; a read-only scan can exceed 50% of it only by reading faster than the copy reads)")
