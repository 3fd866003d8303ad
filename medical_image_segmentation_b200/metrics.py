"""Dataset statistics on the GPU: the mean / std that feed ``Normalize`` in the transform chain.

Mirrors ``compute_mean_and_std(loader)`` of the reference
(medical_image_segmentation/analyze_data/compute_dataset_metrics.py:12-29): iterate a loader whose batches carry
the images in ``batch[0]`` (``[B, C, H, W]``), return per-channel ``(mean, std)`` as float64 tensors, std being the
population standard deviation ``sqrt(E[x^2] - E[x]^2)``.  Inputs here are raw ``torch.uint16`` CUDA batches; the sums
are accumulated exactly (uint64) by ``mis_u16_moments``, so the result does not depend on batch or reduction order.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def accumulate_moments(images: torch.Tensor, sums: torch.Tensor) -> None:
    """sums (uint64-as-int64 ``[C, 2]`` on the device) += per-channel (sum x, sum x^2) of a uint16 ``[B,C,H,W]`` batch."""
    if images.dtype != torch.uint16 or not images.is_cuda:
        raise TypeError("expected a CUDA torch.uint16 batch [B, C, H, W]")
    if images.dim() == 3:
        images = images[:, None]
    images = images.contiguous()
    B, Cc, H, W = images.shape
    with torch.cuda.device(images.device):
        rc = _lib.lib.mis_u16_moments(images.data_ptr(), B, Cc, H * W, sums.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mis_u16_moments")


def compute_mean_and_std(loader, scale: float = 1.0):
    """Per-channel mean and population std over every batch of ``loader`` (``batch[0]`` or the batch itself is the
    uint16 image tensor).  ``scale`` rescales the result (``1/65535`` gives the [0,1] constants ``Normalize`` wants)."""
    sums, n, channels = None, 0, None
    for batch in loader:
        images = batch[0] if isinstance(batch, (tuple, list)) else batch
        if images.dim() == 3:
            images = images[:, None]
        if sums is None:
            channels = images.shape[1]
            sums = torch.zeros((channels, 2), dtype=torch.int64, device=images.device)
        accumulate_moments(images, sums)
        n += images.shape[0] * images.shape[2] * images.shape[3]
    if sums is None:
        raise ValueError("empty loader")
    host = sums.cpu()
    # the device words are unsigned 64-bit; recover values >= 2^63 that int64 shows as negative
    s = [[int(v) & 0xFFFFFFFFFFFFFFFF for v in row] for row in host.tolist()]
    mean = torch.tensor([row[0] / n for row in s], dtype=torch.float64)
    mean_sq = torch.tensor([row[1] / n for row in s], dtype=torch.float64)
    std = torch.sqrt(mean_sq - mean ** 2)
    return mean * scale, std * scale
