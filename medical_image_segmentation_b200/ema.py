"""EMA update of the momentum encoder as one multi-tensor CUDA kernel.

Mirrors ``BYOL.momentum_update(online_encoder, momentum_encoder, m)`` (train/model/byol_pytorch.py:291-296):

    for po, pm in zip(online.parameters(), momentum.parameters()):
        pm.data.mul_(m).add_(po.data, alpha=1.0 - m)

-- two torch kernels per parameter tensor in the reference, one launch over a cached device table here
(csrc/ema.cu; bit-identical arithmetic).  float32 CUDA parameters only; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class _Table:
    def __init__(self, pairs, device):
        entries = np.zeros(len(pairs), _lib.EMA_ENTRY_DTYPE)
        chunk = 0
        for i, (po, pm) in enumerate(pairs):
            entries[i] = (po.data_ptr(), pm.data_ptr(), po.numel(), chunk)
            chunk += int(_lib.lib.mis_ema_chunks(po.numel()))
        self.total_chunks = chunk
        self.n = len(pairs)
        self.dev = torch.from_numpy(entries.view(np.uint8).copy()).to(device)


_tables: dict = {}
launches = 0


@torch.no_grad()
def momentum_update(online_encoder, momentum_encoder, m: float) -> None:
    """``pm = pm * m + po * (1 - m)`` for every parameter pair, in one kernel launch.  ``online_encoder`` /
    ``momentum_encoder`` are modules (as in the reference) or iterables of tensors."""
    global launches
    po_list = list(online_encoder.parameters()) if hasattr(online_encoder, "parameters") else list(online_encoder)
    pm_list = list(momentum_encoder.parameters()) if hasattr(momentum_encoder, "parameters") else list(momentum_encoder)
    if len(po_list) != len(pm_list):
        raise ValueError(f"{len(po_list)} online vs {len(pm_list)} momentum parameters")
    pairs = [(po.data, pm.data) for po, pm in zip(po_list, pm_list) if po.numel() > 0]
    if not pairs:
        return
    device = pairs[0][0].device
    for po, pm in pairs:
        if not (po.is_cuda and pm.is_cuda):
            raise RuntimeError("momentum_update has no CPU path: parameters must be CUDA tensors")
        if po.dtype != torch.float32 or pm.dtype != torch.float32:
            raise TypeError(f"momentum_update expects float32 parameters, got {po.dtype} / {pm.dtype}")
        if po.shape != pm.shape or not (po.is_contiguous() and pm.is_contiguous()) or po.device != device or pm.device != device:
            raise ValueError("parameter pairs must be contiguous, of equal shape and on one device")
    key = tuple((po.data_ptr(), pm.data_ptr(), po.numel()) for po, pm in pairs)
    tab = _tables.get(key)
    if tab is None:
        if len(_tables) > 16:
            _tables.clear()
        tab = _tables[key] = _Table(pairs, device)
    with torch.cuda.device(device):
        rc = _lib.lib.mis_ema_update(tab.dev.data_ptr(), tab.n, tab.total_chunks, float(m), 1.0 - float(m),
                                     C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    _lib.check(rc, "mis_ema_update")
    launches += 1
