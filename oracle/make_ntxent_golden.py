"""Generate tests/golden/ntxent.npz: NT-Xent inputs with loss, per-row log-sum-exp and gradients in float64.

    python -m oracle.make_ntxent_golden

PARITY UNPINNED: the reference contains no NT-Xent (its SSL loss is BYOL, train/model/byol_pytorch.py:181-198), so
these vectors cannot come from it.  They are produced by the lightning-bolts-form restatement
(oracle.loss_oracle.ntxent_loss_bolts / ntxent_rank_sharded_bolts, fp64 autograd) -- written independently of the
cross-entropy form (ntxent_closed_form) the kernels were specified against -- and both forms plus the CUDA kernels are
tested against the committed file.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_oracle as L  # noqa: E402
from tests import synth  # noqa: E402


def main():
    blob = {}
    cases = [("a", 64, 32, False, 0.1), ("b", 128, 128, True, 0.1), ("c", 192, 64, False, 0.07), ("d", 64, 64, True, 0.5)]
    for tag, n, d, clustered, T in cases:
        z1, z2 = synth.embeddings(n, d, seed=100 + n + d, clustered=clustered)
        a = z1.double().requires_grad_(True)
        b = z2.double().requires_grad_(True)
        loss = L.ntxent_loss_bolts(a, b, T)
        loss.backward()
        u = torch.nn.functional.normalize(torch.cat([z1, z2]).double(), dim=1)
        s = u @ u.T / T
        s.fill_diagonal_(float("-inf"))
        blob[f"{tag}_z1"], blob[f"{tag}_z2"] = z1.numpy(), z2.numpy()
        blob[f"{tag}_T"] = T
        blob[f"{tag}_loss"] = float(loss.detach())
        blob[f"{tag}_lse"] = torch.logsumexp(s, dim=1).numpy()
        blob[f"{tag}_dz1"], blob[f"{tag}_dz2"] = a.grad.numpy(), b.grad.numpy()
    # rank-sharded convention, W = 4 simulated ranks
    g = torch.Generator().manual_seed(77)
    z_locals = [torch.randn(2 * 32, 64, generator=g) for _ in range(4)]
    losses, grads = L.ntxent_rank_sharded_bolts(z_locals, 0.1)
    blob["w4_z"] = np.stack([z.numpy() for z in z_locals])
    blob["w4_T"] = 0.1
    blob["w4_loss"] = np.array(losses)
    blob["w4_dz"] = np.stack([gr.numpy() for gr in grads])
    path = os.path.join(ROOT, "tests", "golden", "ntxent.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
