import sys, torch, numpy as np
sys.path.insert(0, ".")
from tests import synth
from medical_image_segmentation_b200.transforms import FusedTwoViewTransforms
H, W, crop, B = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (96, 128, 32, 2)
imgs = synth.batch_512(B, seed=77, H=H, W=W)
import os
t = FusedTwoViewTransforms(crop, (0.2,), (0.2,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), out_dtype=torch.float32, use_tma=int(os.environ.get("TMA", "0")))
torch.manual_seed(31)
t(torch.from_numpy(imgs).cuda())
torch.cuda.synchronize()
print("ok", float(t.views_buffer.abs().mean()))
