"""CPU oracle for the contrastive objective (TEST INFRASTRUCTURE ONLY).

``byol_cosine_loss``  restates BYOL.cosine_similarity_loss
                      (train/model/byol_pytorch.py:181-198) -- pinned against the
                      reference by tests/golden/byol_loss.npz.
``ntxent_*``          canonical SimCLR NT-Xent.  PARITY
                      UNPINNED: the reference has no NT-Xent; these are
                      checked against torch autograd in fp64 only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def byol_cosine_loss(preds: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """2 - 2 * mean_i <p_i/|p_i|, t_i/|t_i|>   (byol_pytorch.py:196-198)."""
    p = F.normalize(preds, dim=-1, p=2)
    t = F.normalize(targets, dim=-1, p=2)
    return 2 - 2 * (p * t).sum(dim=-1).mean()


def ntxent_loss(z_a: torch.Tensor, z_b: torch.Tensor, temperature: float = 0.1) -> torch.Tensor:
    """A.4: z=[z_a;z_b]; u=normalize(z); S=u u^T/T, diag=-inf; mean CE against (i+N) mod 2N."""
    z = torch.cat([z_a, z_b], dim=0)
    n2 = z.shape[0]
    n = n2 // 2
    u = F.normalize(z, dim=1)
    s = (u @ u.T) / temperature
    s = s.masked_fill(torch.eye(n2, dtype=torch.bool), float("-inf"))
    target = (torch.arange(n2) + n) % n2
    return F.cross_entropy(s, target)


def ntxent_closed_form(z_a: np.ndarray, z_b: np.ndarray, temperature: float = 0.1):
    """Loss, per-row lse and dL/dz in float64 without autograd (A.4 closed form).

    P = exp(S - lse) (P_ii = 0); G = (P - onehot(p)) / 2N; dU = (G + G^T) U / T;
    dz_i = (dU_i - u_i <u_i, dU_i>) / |z_i|.
    """
    z = np.concatenate([np.asarray(z_a, np.float64), np.asarray(z_b, np.float64)], axis=0)
    n2 = z.shape[0]
    n = n2 // 2
    norm = np.maximum(np.sqrt((z * z).sum(1, keepdims=True)), 1e-12)
    u = z / norm
    s = (u @ u.T) / temperature
    np.fill_diagonal(s, -np.inf)
    m = s.max(1, keepdims=True)
    lse = (m + np.log(np.exp(s - m).sum(1, keepdims=True)))[:, 0]
    pos = (np.arange(n2) + n) % n2
    loss = float((lse - s[np.arange(n2), pos]).mean())
    p = np.exp(s - lse[:, None])
    g = p.copy()
    g[np.arange(n2), pos] -= 1.0
    g /= n2
    du = (g + g.T) @ u / temperature
    dz = (du - u * (u * du).sum(1, keepdims=True)) / norm
    return loss, lse, dz[:n], dz[n:]


def ntxent_rank_sharded(z_locals: list[torch.Tensor], temperature: float = 0.1):
    """A.5 distributed convention, simulated in one process with autograd (fp64).

    ``z_locals[r]`` is rank r's ``[2*B_local, D]`` block laid out ``[v1_local; v2_local]``;
    the global matrix is the rank-major concatenation.  Rank r's loss L_r is the mean over
    its own rows against all columns; the gradient each rank must return is
    ``sum_r' dL_r'/dZ_local`` (= W * dL_global/dZ_local), so that DDP's 1/W averaging of
    parameter gradients reproduces the global gradient.
    Returns (list of L_r, list of dZ_local).
    """
    zs = [z.detach().double().clone().requires_grad_(True) for z in z_locals]
    z_all = torch.cat(zs, dim=0)
    n2 = z_all.shape[0]
    u = F.normalize(z_all, dim=1)
    s = (u @ u.T) / temperature
    s = s.masked_fill(torch.eye(n2, dtype=torch.bool), float("-inf"))
    losses, row0 = [], 0
    pos = torch.empty(n2, dtype=torch.long)
    for z in zs:
        rows = z.shape[0]
        half = rows // 2
        idx = torch.arange(rows)
        pos[row0:row0 + rows] = row0 + (idx + half) % rows
        row0 += rows
    row0 = 0
    for z in zs:
        rows = z.shape[0]
        sl = slice(row0, row0 + rows)
        losses.append(F.cross_entropy(s[sl], pos[sl]))
        row0 += rows
    torch.stack(losses).sum().backward()
    return [float(l) for l in losses], [z.grad.clone() for z in zs]
