"""CPU oracle for the contrastive objective (TEST INFRASTRUCTURE ONLY).

``byol_cosine_loss``  restates BYOL.cosine_similarity_loss
                      (train/model/byol_pytorch.py:181-198) -- pinned against the
                      reference by tests/golden/byol_loss.npz.
``ntxent_*``          canonical SimCLR NT-Xent.  PARITY
                      UNPINNED: the reference has no NT-Xent; these are
                      checked against torch autograd in fp64 only.
``ntxent_*_bolts``    a SECOND, independently written restatement: the exp-sum form of the
                      lightning-bolts SimCLR module the reference's callbacks / optimiser were
                      adapted from (pl_bolts @748715e, models/self_supervised/simclr/simclr_module.py
                      ``nt_xent_loss``; source not under /root/reference, restated from the
                      published algorithm): no -inf mask, no log-sum-exp -- the self term is
                      removed by subtracting e^(1/T), with eps = 1e-6 clamps.  The two forms, the
                      committed vectors (tests/golden/ntxent.npz, oracle/make_ntxent_golden.py) and
                      the CUDA kernels are tested against each other; that removes single-author
                      risk, it does not pin the loss to the reference (nothing can).
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


def ntxent_loss_bolts(z_a: torch.Tensor, z_b: torch.Tensor, temperature: float = 0.1, eps: float = 1e-6) -> torch.Tensor:
    """lightning-bolts SimCLR ``nt_xent_loss`` (single process): out = normalize(z);
    cov = out out^T; neg_i = sum_j exp(cov_ij / T) - e^(1/T) (clamped at eps);
    pos_i = exp(<out1_i, out2_i> / T); loss = -mean log(pos / (neg + eps))."""
    import math
    out_1 = F.normalize(z_a, dim=1)
    out_2 = F.normalize(z_b, dim=1)
    out = torch.cat([out_1, out_2], dim=0)
    cov = torch.mm(out, out.t().contiguous())
    sim = torch.exp(cov / temperature)
    neg = sim.sum(dim=-1)
    row_sub = torch.full_like(neg, math.e ** (1 / temperature))
    neg = torch.clamp(neg - row_sub, min=eps)
    pos = torch.exp(torch.sum(out_1 * out_2, dim=-1) / temperature)
    pos = torch.cat([pos, pos], dim=0)
    return -torch.log(pos / (neg + eps)).mean()


def ntxent_rank_sharded_bolts(z_locals: list[torch.Tensor], temperature: float = 0.1, eps: float = 1e-6):
    """The bolts form with its SyncFunction gather, simulated in one process (fp64 autograd): every rank r computes
    -mean log(pos / (neg + eps)) over ITS rows [z1_r; z2_r] against out_dist = [z1_all; z2_all] (bolts' canonical
    gathered layout, not the rank-major one of the product -- the loss is invariant to that permutation), and
    SyncFunction's backward hands each rank the SUM over ranks of dL_r'/dz_local.  Returns (losses, dz_locals)."""
    import math
    zs = [z.detach().double().clone().requires_grad_(True) for z in z_locals]
    halves = [(F.normalize(z[:z.shape[0] // 2], dim=1), F.normalize(z[z.shape[0] // 2:], dim=1)) for z in zs]
    out_dist = torch.cat([h[0] for h in halves] + [h[1] for h in halves], dim=0)
    losses = []
    for o1, o2 in halves:
        out = torch.cat([o1, o2], dim=0)
        sim = torch.exp(torch.mm(out, out_dist.t().contiguous()) / temperature)
        neg = torch.clamp(sim.sum(dim=-1) - math.e ** (1 / temperature), min=eps)
        pos = torch.exp(torch.sum(o1 * o2, dim=-1) / temperature)
        pos = torch.cat([pos, pos], dim=0)
        losses.append(-torch.log(pos / (neg + eps)).mean())
    torch.stack(losses).sum().backward()
    return [float(l.detach()) for l in losses], [z.grad.clone() for z in zs]
