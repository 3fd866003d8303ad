"""cfg5 sweep: K1 alone over crop sizes / batch sizes / kernel variants.  python scripts/sweep_aug.py [json_out]"""
import json, sys
import torch
sys.path.insert(0, ".")
from medical_image_segmentation_b200.transforms import FusedTwoViewTransforms, algorithmic_bytes

PEAK = 6548.8
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
rows = []
x_all = torch.randint(0, 65536, (4096, 1, 512, 512), dtype=torch.int32, device="cuda").to(torch.uint16)
for crop in (96, 224, 256):
    for B in (256, 1024, 4096):
        for variant in (0, 3):
            x = x_all[:B]
            t = FusedTwoViewTransforms(crop, (0.227358,), (0.237160,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0), use_tma=variant)
            torch.manual_seed(0)
            params = t.to_view_major(t.draw_params(B, 512, 512))
            nbytes = algorithmic_bytes(params, 1, crop)
            out = torch.empty((2 * B, 1, crop, crop), dtype=torch.bfloat16, device="cuda")
            for _ in range(3):
                t.apply(x, params, out)
            torch.cuda.synchronize()
            n = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                t.apply(x, params, out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            r = dict(crop=crop, slices=B, variant=variant, ms=round(ms, 4), views_per_s=round(2 * B / ms * 1e3),
                     gbs=round(nbytes / ms / 1e6, 1), frac=round(nbytes / ms / 1e6 / PEAK, 4))
            rows.append(r)
            print(r, flush=True)
if len(sys.argv) > 1:
    json.dump({"peak_gbs": PEAK, "rows": rows}, open(sys.argv[1], "w"), indent=1)
