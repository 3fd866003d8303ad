"""Small fixed workload for ncu: python scripts/prof_aug.py [B] [crop] [use_tma]."""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_segmentation_b200.transforms import FusedTwoViewTransforms

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 224
use_tma = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randint(0, 65536, (B, 1, 512, 512), dtype=torch.int32, device="cuda", generator=g).to(torch.uint16)
t = FusedTwoViewTransforms(crop, (0.227358,), (0.237160,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), use_tma=use_tma)
torch.manual_seed(0)
params = t.to_view_major(t.draw_params(B, 512, 512))
out = torch.empty((2 * B, 1, crop, crop), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    t.apply(x, params, out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
