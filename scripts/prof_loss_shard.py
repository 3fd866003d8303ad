"""One rank's share of the cfg3 NT-Xent (rows of one GPU against the gathered matrix of all), on one GPU through the
ABI: python scripts/prof_loss_shard.py [rows] [cols] [D] [iters].  Used for ncu (tensor-pipe activity at the shard shape)."""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_segmentation_b200.loss import CudaKernels

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
g = torch.Generator(device="cuda").manual_seed(0)
z_all = torch.randn(cols, D, device="cuda", generator=g)
_, u_all, rinv_all = CudaKernels.prep(z_all)
z, rinv = z_all[:rows].contiguous(), rinv_all[:rows].contiguous()
scratch = CudaKernels.scratch(rows, cols, D, "cuda")
# the other ranks' lse: one forward per row block (values only matter for realism of the backward's exponentials)
lse_all = torch.cat([CudaKernels.fwd(u_all, r0, rows, 10.0, scratch)[0] for r0 in range(0, cols, rows)])
one = torch.ones(1, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(iters + 1):
    if it == 1:
        e0.record()
    CudaKernels.fwd(u_all, 0, rows, 10.0, scratch)
    CudaKernels.bwd(u_all, lse_all, z, rinv, 0, 10.0, one, scratch)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"rows={rows} cols={cols} D={D}: {ms:.4f} ms per fwd+bwd (kernels only), {6.0*rows*cols*D/ms/1e9:.1f} TFLOP/s algorithmic")
