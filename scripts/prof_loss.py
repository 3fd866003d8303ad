"""Fixed NT-Xent workload for ncu / timing: python scripts/prof_loss.py [N] [D] [iters]."""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_segmentation_b200 import nt_xent_rows
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn(2 * N, D, device="cuda", generator=g).requires_grad_(True)
for _ in range(iters):
    z.grad = None
    nt_xent_rows(z, 0.1).backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    z.grad = None
    nt_xent_rows(z, 0.1).backward()
e1.record(); torch.cuda.synchronize()
print(f"2N={2*N} D={D}: {e0.elapsed_time(e1)/iters:.4f} ms per fwd+bwd, {6*(2*N)**2*D/(e0.elapsed_time(e1)/iters)/1e9:.1f} TFLOP/s")
