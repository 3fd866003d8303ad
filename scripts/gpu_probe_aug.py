"""First-contact probe for K1: python scripts/gpu_probe_aug.py <use_tma 0|1>."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import aug_oracle as A
from tests import synth
from medical_image_segmentation_b200.transforms import FusedTwoViewTransforms, algorithmic_bytes

use_tma = int(sys.argv[1])
MEAN, STD = 0.227358, 0.237160
for (H, W, crop, B) in ((96, 128, 32, 4), (512, 512, 224, 4), (512, 512, 96, 4), (512, 512, 256, 2), (256, 768, 112, 2)):
    imgs = synth.batch_512(B, seed=77, H=H, W=W)
    t = FusedTwoViewTransforms(crop, (MEAN,), (STD,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), out_dtype=torch.float32, use_tma=use_tma)
    torch.manual_seed(31)
    t(torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    out = t.views_buffer.cpu().numpy()
    p = t.last_params
    worst = 0
    for i in range(B):
        for v in range(2):
            r = p[2 * i + v]
            pr = dict(top=int(r["top"]), left=int(r["left"]), h=int(r["h"]), w=int(r["w"]), flip=bool(r["flags"] & 1),
                      jitter=bool(r["flags"] & 2), order=tuple(int(x) for x in r["order"]),
                      brightness=float(r["brightness"]), contrast=float(r["contrast"]))
            ref = A.apply_view(imgs[i], pr, crop, MEAN, STD)
            e = np.abs(out[v * B + i, 0] - ref).max()
            worst = max(worst, e)
    print(f"tma={use_tma} {H}x{W} -> {crop}: max abs err {worst:.3e}", flush=True)

# timing at the bench config
B, crop = 1024, 224
x = torch.randint(0, 65536, (B, 1, 512, 512), dtype=torch.int32, device="cuda").to(torch.uint16)
t = FusedTwoViewTransforms(crop, (MEAN,), (STD,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), use_tma=use_tma)
torch.manual_seed(0)
params = t.to_view_major(t.draw_params(B, 512, 512))
nbytes = algorithmic_bytes(params, 1, crop)
out = torch.empty((2 * B, 1, crop, crop), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    t.apply(x, params, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    t.apply(x, params, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"tma={use_tma} B={B} crop={crop}: {ms:.3f} ms/launch, {2*B/ms*1e3:.3e} views/s, {nbytes/ms/1e6:.1f} GB/s algorithmic "
      f"({nbytes/ms/1e6/6548.8*100:.1f}% of 6548.8)")
