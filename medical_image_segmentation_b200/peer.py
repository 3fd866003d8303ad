"""Symmetric (peer-mapped) buffers behind the NVLink exchange of the NT-Xent loss.

The reference gathers embeddings with ``pl_module.all_gather`` (``concat_all_gather``, train/callback/knn.py:143-144).
Here the two gathers of the loss -- normalised rows (forward) and log-sum-exp scalars (backward) -- are written by
the kernels that produce them straight into the same buffer of every GPU of the node over NVLink (csrc/ntxent.cu,
``Peers``), so no collective call sits on the data path.  ``torch.distributed._symmetric_memory`` is used only to
allocate the buffer and exchange the peer pointers (plumbing); when it is unavailable the loss falls back to
``torch.distributed.all_gather_into_tensor`` (NCCL).
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

import torch
import torch.distributed as dist

MAX_PEERS = 8
_FLAG_BYTES = 256            # uint32 [2][8] flags + 2 producer counters + timeout word, padded


def _al256(n: int) -> int:
    return (n + 255) // 256 * 256


class PeerExchange:
    """One symmetric allocation per (group, rows, D): flags | u_all[2] | lse_all[2] (double-buffered by epoch parity)."""

    def __init__(self, group, rows: int, D: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > MAX_PEERS:
            raise RuntimeError(f"peer exchange covers one node (<= {MAX_PEERS} ranks), got {self.world}")
        self.rows, self.D = rows, D
        self.cols = self.world * rows
        u_bytes = _al256(self.cols * D * 4)
        l_bytes = _al256(self.cols * 4)
        self.off_u = [_FLAG_BYTES, _FLAG_BYTES + u_bytes]
        self.off_l = [_FLAG_BYTES + 2 * u_bytes, _FLAG_BYTES + 2 * u_bytes + l_bytes]
        nbytes = _FLAG_BYTES + 2 * u_bytes + 2 * l_bytes
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        bases = [int(p) for p in self.hdl.buffer_ptrs]
        if len(bases) != self.world or bases[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric memory handle does not describe this buffer")
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)                              # nobody signals before every flag block is zero
        arr = C.c_void_p * MAX_PEERS
        pad = [0] * (MAX_PEERS - self.world)
        self.flag_peers = arr(*(bases + pad))
        self.u_peers = [arr(*([b + o for b in bases] + pad)) for o in self.off_u]
        self.l_peers = [arr(*([b + o for b in bases] + pad)) for o in self.off_l]
        self.flags_ptr = self.buf.data_ptr()
        self.u_all = [self.buf[o:o + self.cols * D * 4].view(torch.float32).view(self.cols, D) for o in self.off_u]
        self.lse_all = [self.buf[o:o + self.cols * 4].view(torch.float32) for o in self.off_l]
        self.epoch = 0
        self.scratch = None                              # kernel scratch + rinv per parity (allocated by loss.py)
        self.rinv = None
        self.in_flight = [0, 0]                          # forwards whose backward has not run yet, per parity

    def timed_out(self) -> bool:
        """True if a consumer gave up waiting for a peer (synchronises the device)."""
        return bool(self.buf[4 * (2 * MAX_PEERS + 2):4 * (2 * MAX_PEERS + 3)].view(torch.int32).item())


_cache: dict = {}
_disabled_reason: str | None = None


def mode() -> str:
    return os.environ.get("MIS_NTXENT_EXCHANGE", "auto").lower()


def get_exchange(group, rows: int, D: int, device: torch.device):
    """The cached exchange for this shape, or None when peer memory is unavailable / disabled (then NCCL is used)."""
    global _disabled_reason
    if mode() == "nccl" or _disabled_reason is not None:
        return None
    key = (id(group), rows, D, device.index)
    ex = _cache.get(key)
    if ex is None:
        try:
            ex = PeerExchange(group, rows, D, device)
        except Exception as e:                            # every rank takes the same branch: set-up is collective
            _disabled_reason = f"{type(e).__name__}: {e}"
            if mode() == "peer":
                raise
            warnings.warn(f"NT-Xent peer-memory exchange unavailable ({_disabled_reason}); using NCCL all-gather")
            return None
        _cache[key] = ex
    return ex


def check_timeouts() -> None:
    """Raise if any consumer gave up waiting for a peer's flag (synchronises the device; call it off the hot path)."""
    for ex in _cache.values():
        if ex.timed_out():
            raise RuntimeError(f"NT-Xent peer exchange: rank {ex.rank} timed out waiting for a peer's flag "
                               "(a rank died or fell more than ~2 s behind); results of that step are invalid")


check_health = check_timeouts
