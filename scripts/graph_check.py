import os, sys, torch
sys.path.insert(0, ".")
from oracle import loss_oracle as L
from medical_image_segmentation_b200 import nt_xent_loss
import numpy as np
for mode in ("0", "1"):
    os.environ["MIS_NTXENT_GRAPH"] = mode
    for rep in range(3):
        g = torch.Generator().manual_seed(rep)
        z1, z2 = torch.randn(256, 128, generator=g), torch.randn(256, 128, generator=g)
        a, b = z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True)
        loss = nt_xent_loss(a, b, 0.1); (2.0 * loss).backward()
        ra, rb = z1.double().requires_grad_(True), z2.double().requires_grad_(True)
        rl = L.ntxent_loss(ra, rb, 0.1); (2.0 * rl).backward()
        fro = (torch.cat([a.grad, b.grad]).cpu().double() - torch.cat([ra.grad, rb.grad])).norm() / torch.cat([ra.grad, rb.grad]).norm()
        print(mode, rep, float(loss), float(rl), float(fro))
        assert abs(float(loss) - float(rl)) < 1e-3 * abs(float(rl)) and fro < 1e-3
    # two forwards before the backwards (graph busy -> eager fallback)
    a, b = z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True)
    l1 = nt_xent_loss(a, b, 0.1); l2 = nt_xent_loss(a, b, 0.1); (l1 + l2).backward()
    print("double", float(l1), float(l2))
print("GRAPH TEST OK")
