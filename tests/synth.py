"""Deterministic synthetic inputs: uint16 slices and embeddings.

Used by the tests, bench.py and oracle/make_golden.py so that every box regenerates the
same bytes from a seed (torch.Generator CPU streams are stable for a fixed torch build).
"""
from __future__ import annotations

import numpy as np
import torch


def uniform_slice(H: int, W: int, seed: int) -> np.ndarray:
    """i.i.d. uniform 0..65535 -- worst case for caches and clamps."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.randint(0, 65536, (H, W), generator=g, dtype=torch.int32).numpy().astype(np.uint16)


def ct_like_slice(H: int, W: int, seed: int) -> np.ndarray:
    """Sum of 8 low-frequency cosines + 2 % noise, min-max scaled to exactly 0..65535.

    Imitates the value distribution written by analyze_data/create_subset.py:215-222
    (every slice spans the full uint16 range).
    """
    g = torch.Generator().manual_seed(int(seed))
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H, dtype=torch.float64),
                            torch.linspace(0, 1, W, dtype=torch.float64), indexing="ij")
    img = torch.zeros(H, W, dtype=torch.float64)
    for _ in range(8):
        fy, fx, ph, amp = torch.rand(4, generator=g, dtype=torch.float64)
        img += (0.3 + amp) * torch.cos(2 * torch.pi * (4 * fy * yy + 4 * fx * xx + ph))
    img += 0.02 * (img.max() - img.min()) * torch.rand(H, W, generator=g, dtype=torch.float64)
    img = (img - img.min()) / (img.max() - img.min())
    return torch.round(img * 65535.0).to(torch.int32).numpy().astype(np.uint16)


def batch_512(n: int, seed: int = 1234, H: int = 512, W: int = 512) -> np.ndarray:
    """[n,H,W] uint16: even indices CT-like, odd indices uniform noise."""
    return np.stack([ct_like_slice(H, W, seed + i) if i % 2 == 0 else uniform_slice(H, W, seed + i)
                     for i in range(n)])


def embeddings(n: int, d: int, seed: int = 0, clustered: bool = False, dtype=torch.float32):
    """z1, z2 = randn(n,d); clustered: z2 = z1 + 0.3*randn (sharp softmax)."""
    g = torch.Generator().manual_seed(int(seed))
    z1 = torch.randn(n, d, generator=g, dtype=torch.float32)
    z2 = torch.randn(n, d, generator=g, dtype=torch.float32)
    if clustered:
        z2 = z1 + 0.3 * z2
    return z1.to(dtype), z2.to(dtype)
