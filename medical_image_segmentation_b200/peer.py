"""Symmetric (peer-mapped) buffers behind the NVLink exchange of the NT-Xent loss.

The reference gathers embeddings with ``pl_module.all_gather`` (``concat_all_gather``, train/callback/knn.py:143-144).
Here the two gathers of the loss -- normalised rows (forward) and log-sum-exp scalars (backward) -- are written by
the kernels that produce them straight into the same buffer of every GPU of the node over NVLink, and the kernels that
consume them wait for the flags themselves (csrc/ntxent.cu, ``PeerCtl``), so no collective call sits on the data path.
``torch.distributed._symmetric_memory`` is used only to allocate the buffer and exchange the peer pointers (plumbing).
Whether the peer path is used is decided COLLECTIVELY (an all-reduce over the set-up outcome of every rank): either all
ranks of the group use it or all use ``torch.distributed.all_gather_into_tensor`` (NCCL).

A consumer waits ``MIS_PEER_TIMEOUT_S`` seconds (default 600, the order of the NCCL watchdog) for a peer; after that
the kernel traps and the step fails with a CUDA error -- it never continues on stale rows.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

import torch
import torch.distributed as dist

MAX_PEERS = 8
_CTL_BYTES = 256             # PeerCtl: uint32 flag[2][8], producer counter, epoch, abort word, padded (csrc/ntxent.cu)
_OFF_EPOCH = 4 * (2 * MAX_PEERS + 1)
_OFF_ABORT = 4 * (2 * MAX_PEERS + 2)


def _al256(n: int) -> int:
    return (n + 255) // 256 * 256


def timeout_s() -> float:
    return float(os.environ.get("MIS_PEER_TIMEOUT_S", "600"))


class PeerExchange:
    """One symmetric allocation per (group, rows, D): ctl | u_all[2] | lse_all[2] (double-buffered by epoch parity)."""

    def __init__(self, group, rows: int, D: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > MAX_PEERS:
            raise RuntimeError(f"peer exchange covers one node (<= {MAX_PEERS} ranks), got {self.world}")
        from . import _lib
        self.rows, self.D = rows, D
        self.rows_pad = int(_lib.lib.mis_ntxent_padded_rows(rows))     # every rank's block is padded to 128-row tiles
        self.cols = self.world * self.rows_pad
        u_bytes = _al256(self.cols * D * 4)
        l_bytes = _al256(self.cols * 4)
        self.off_u = [_CTL_BYTES, _CTL_BYTES + u_bytes]
        self.off_l = [_CTL_BYTES + 2 * u_bytes, _CTL_BYTES + 2 * u_bytes + l_bytes]
        nbytes = _CTL_BYTES + 2 * u_bytes + 2 * l_bytes
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        bases = [int(p) for p in self.hdl.buffer_ptrs]
        if len(bases) != self.world or bases[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric memory handle does not describe this buffer")
        self.buf.zero_()
        torch.cuda.synchronize(device)
        arr = C.c_void_p * MAX_PEERS
        pad = [0] * (MAX_PEERS - self.world)
        self.ctl_peers = arr(*(bases + pad))
        self.u_peers = [arr(*([b + o for b in bases] + pad)) for o in self.off_u]
        self.l_peers = [arr(*([b + o for b in bases] + pad)) for o in self.off_l]
        self.scratch = None                              # kernel scratch, rinv (allocated by loss.py)
        self.rinv = None
        self.graph = None                                # captured evaluation (MIS_NTXENT_GRAPH=1)

    def device_epoch(self) -> int:
        """Forwards completed on this rank (synchronises the device)."""
        return int(self.buf[_OFF_EPOCH:_OFF_EPOCH + 4].view(torch.int32).item())

    def aborted(self) -> bool:
        """True if a consumer gave up waiting for a peer (synchronises the device)."""
        return bool(self.buf[_OFF_ABORT:_OFF_ABORT + 4].view(torch.int32).item())


_cache: dict = {}
_disabled_reason: str | None = None


def mode() -> str:
    return os.environ.get("MIS_NTXENT_EXCHANGE", "auto").lower()


def get_exchange(group, rows: int, D: int, device: torch.device):
    """The cached exchange for this shape, or None when peer memory is unavailable / disabled (then NCCL is used).

    Set-up is collective and so is its outcome: every rank reports success or failure, the minimum is all-reduced, and
    all ranks take the same transport.  (A rank-local failure would otherwise leave that rank in all_gather while the
    others spin on flags.)"""
    global _disabled_reason
    if mode() == "nccl" or _disabled_reason is not None:
        return None
    key = (id(group), rows, D, device.index)
    ex = _cache.get(key)
    if ex is None:
        err = None
        try:
            ex = PeerExchange(group, rows, D, device)
        except Exception as e:
            err, ex = f"{type(e).__name__}: {e}", None
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)     # also the barrier: every control block is zero
        if int(ok.item()) == 0:
            _disabled_reason = err or "set-up failed on another rank"
            if mode() == "peer":
                raise RuntimeError(f"NT-Xent peer-memory exchange unavailable: {_disabled_reason}")
            warnings.warn(f"NT-Xent peer-memory exchange unavailable ({_disabled_reason}); all ranks use NCCL all-gather")
            return None
        _cache[key] = ex
    return ex


def check_health() -> None:
    """Raise if any consumer gave up waiting for a peer's flag (synchronises the device; call it off the hot path).
    A time-out also traps the kernel, so normally the CUDA error surfaces first."""
    for ex in _cache.values():
        if ex.aborted():
            raise RuntimeError(f"NT-Xent peer exchange: rank {ex.rank} timed out waiting for a peer's flag "
                               f"(a rank died or fell more than {timeout_s():.0f} s behind)")


check_timeouts = check_health
