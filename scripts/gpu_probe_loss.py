"""First-contact probe for K2/K3."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import loss_oracle as L
from tests import synth
from medical_image_segmentation_b200 import nt_xent_loss

for (n, d, cl) in ((64, 32, False), (128, 128, False), (1024, 128, True)):
    z1, z2 = synth.embeddings(n, d, seed=1, clustered=cl)
    a = z1.cuda().requires_grad_(True); b = z2.cuda().requires_grad_(True)
    loss = nt_xent_loss(a, b, 0.1)
    torch.cuda.synchronize()
    print("fwd ok", float(loss), flush=True)
    loss.backward()
    torch.cuda.synchronize()
    rl, _, d1, d2 = L.ntxent_closed_form(z1.numpy(), z2.numpy(), 0.1)
    g = a.grad.cpu().numpy().astype(np.float64)
    print(f"n={n} d={d} clustered={cl}: loss {float(loss):.6f} ref {rl:.6f} rel {abs(float(loss)-rl)/rl:.2e}  "
          f"grad fro rel {np.linalg.norm(g-d1)/np.linalg.norm(d1):.3e} max rel {np.abs(g-d1).max()/np.abs(d1).max():.3e}", flush=True)
# timing cfg2
z1, z2 = synth.embeddings(1024, 128, seed=0)
a = z1.cuda().requires_grad_(True); b = z2.cuda().requires_grad_(True)
for _ in range(3):
    a.grad = b.grad = None
    nt_xent_loss(a, b, 0.1).backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    a.grad = b.grad = None
    nt_xent_loss(a, b, 0.1).backward()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"NT-Xent fwd+bwd 2N=2048 D=128: {ms:.4f} ms  -> {6*2048*2048*128/ms/1e9:.1f} TFLOP/s algorithmic")
