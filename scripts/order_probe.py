"""Does the order of the views inside a K1 launch matter?  Time the same table as drawn and sorted by descending crop area."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from medical_image_segmentation_b200 import FusedTwoViewTransforms

for B in (256, 512, 1024, 4096):
    x = torch.randint(0, 65536, (B, 1, 512, 512), dtype=torch.int32, device="cuda").to(torch.uint16)
    t = FusedTwoViewTransforms(224, (0.227,), (0.237,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0))
    torch.manual_seed(7)
    p = t.to_view_major(t.draw_params(B, 512, 512))
    out = torch.empty((2 * B, 1, 224, 224), dtype=torch.bfloat16, device="cuda")
    area = p["h"].astype(np.int64) * p["w"]
    tables = {"as drawn": p, "descending area": p[np.argsort(-area, kind="stable")], "ascending area": p[np.argsort(area, kind="stable")]}
    for name, tab in tables.items():
        tab = np.ascontiguousarray(tab)
        for _ in range(5):
            t.apply(x, tab, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50):
            t.apply(x, tab, out)
        e1.record(); torch.cuda.synchronize()
        print(f"B={B:5d} {name:16s}: {e0.elapsed_time(e1) / 50:.4f} ms", flush=True)
