"""World-size-2 gloo test of the multi-rank NT-Xent host path (all-gather layout, option-L backward).

The device kernels are replaced by tests/cpu_kernels.TorchKernels; everything else is the product code in
medical_image_segmentation_b200/loss.py.  Oracle: oracle.loss_oracle.ntxent_rank_sharded (fp64 autograd)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, D, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from medical_image_segmentation_b200.loss import nt_xent_rows
    from tests.cpu_kernels import TorchKernels
    g = torch.Generator().manual_seed(3)
    z_locals = [torch.randn(2 * B, D, generator=g, dtype=torch.float64) for _ in range(world)]
    z = z_locals[rank].clone().requires_grad_(True)
    loss = nt_xent_rows(z, 0.1, dist.group.WORLD, _kernels=TorchKernels)
    (2.0 * loss).backward()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss.item(), grad=z.grad.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_matches_sharded_oracle(tmp_path):
    from oracle import loss_oracle as L
    world, B, D = 2, 6, 16
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, D, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(3)
    z_locals = [torch.randn(2 * B, D, generator=g, dtype=torch.float64) for _ in range(world)]
    ref_losses, ref_grads = L.ntxent_rank_sharded(z_locals, 0.1)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert abs(float(got["loss"]) - ref_losses[r]) < 1e-12
        assert np.abs(got["grad"] - 2.0 * ref_grads[r].numpy()).max() < 1e-12


def test_single_process_path_without_group():
    from medical_image_segmentation_b200.loss import nt_xent_rows
    from oracle import loss_oracle as L
    from tests.cpu_kernels import TorchKernels
    g = torch.Generator().manual_seed(5)
    z = torch.randn(24, 8, generator=g, dtype=torch.float64).requires_grad_(True)
    loss = nt_xent_rows(z, 0.2, None, _kernels=TorchKernels)
    loss.backward()
    ref, _, d1, d2 = L.ntxent_closed_form(z.detach()[:12].numpy(), z.detach()[12:].numpy(), 0.2)
    assert abs(loss.item() - ref) < 1e-12
    assert np.abs(z.grad.numpy() - np.concatenate([d1, d2])).max() < 1e-13
