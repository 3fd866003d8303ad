"""Multi-GPU parity of the NT-Xent path (NCCL all-gather + option-L backward) against the sharded oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from oracle import loss_oracle as L
from medical_image_segmentation_b200 import nt_xent_loss

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (B, D, T) in ((64, 64, 0.1), (256, 128, 0.1), (128, 256, 0.2), (256, 128, 0.1), (256, 128, 0.1), (96, 64, 0.1), (96, 64, 0.1), (200, 128, 0.1)):
    g = torch.Generator().manual_seed(11)
    z_locals = [torch.randn(2 * B, D, generator=g) for _ in range(world)]
    ref_losses, ref_grads = L.ntxent_rank_sharded(z_locals, T)
    z = z_locals[rank].cuda()
    a = z[:B].clone().requires_grad_(True)
    b = z[B:].clone().requires_grad_(True)
    loss = nt_xent_loss(a, b, T)          # default group
    loss.backward()
    got = torch.cat([a.grad, b.grad]).cpu().double().numpy()
    ref = ref_grads[rank].numpy()
    fro = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    lrel = abs(loss.item() - ref_losses[rank]) / abs(ref_losses[rank])
    good = fro <= 1e-3 and lrel <= 1e-3
    ok &= good
    print(f"rank {rank}/{world} B={B} D={D} T={T}: loss rel {lrel:.2e} grad fro rel {fro:.2e} {'OK' if good else 'FAIL'}", flush=True)
# evaluation patterns: a forward under no_grad between two training evaluations, a scaled loss, and two forwards whose
# backwards run later in one autograd pass (each evaluation issues its own backward kernels right behind its forward)
B, D, T = 128, 128, 0.1
g = torch.Generator().manual_seed(23)
za = [torch.randn(2 * B, D, generator=g) for _ in range(world)]
zb = [torch.randn(2 * B, D, generator=g) for _ in range(world)]
la, ga = L.ntxent_rank_sharded(za, T)
lb, gb = L.ntxent_rank_sharded(zb, T)
from medical_image_segmentation_b200 import nt_xent_rows
for it in range(3):
    xa = za[rank].cuda().requires_grad_(True)
    xb = zb[rank].cuda().requires_grad_(True)
    with torch.no_grad():
        l0 = nt_xent_rows(xb, T, dist.group.WORLD)
    l1 = nt_xent_rows(xa, T, dist.group.WORLD)
    l2 = nt_xent_rows(xb, T, dist.group.WORLD)
    (3.0 * l1 + l2).backward()
    e = [abs(l0.item() - lb[rank]) / abs(lb[rank]), abs(l1.item() - la[rank]) / abs(la[rank]), abs(l2.item() - lb[rank]) / abs(lb[rank]),
         np.linalg.norm(xa.grad.cpu().double().numpy() - 3.0 * ga[rank].numpy()) / np.linalg.norm(3.0 * ga[rank].numpy()),
         np.linalg.norm(xb.grad.cpu().double().numpy() - gb[rank].numpy()) / np.linalg.norm(gb[rank].numpy())]
    good = max(e) <= 1e-3
    ok &= good
    print(f"rank {rank}/{world} patterns pass {it}: no_grad loss {e[0]:.1e}, losses {e[1]:.1e} {e[2]:.1e}, grads {e[3]:.1e} {e[4]:.1e} "
          f"{'OK' if good else 'FAIL'}", flush=True)
from medical_image_segmentation_b200 import peer
print(f"rank {rank}: exchange mode {peer.mode()}, peer exchanges in use: {len(peer._cache)}, disabled: {peer._disabled_reason}", flush=True)
try:
    peer.check_health()
except RuntimeError as e:
    ok = False
    print(f"rank {rank}: {e}", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK", "PASS" if flag.item() == 1.0 else "FAIL")
sys.exit(0 if flag.item() == 1.0 else 1)
