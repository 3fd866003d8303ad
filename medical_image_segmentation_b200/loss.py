"""Contrastive objectives behind the reference's loss call surface.

The reference calls ``self.cosine_similarity_loss(preds, targets)`` from ``BYOL.training_step``
(train/model/byol_pytorch.py:181-198, called at :217).  This module provides, with the same
``(a, b) -> scalar`` shape,

* ``nt_xent_loss(z_a, z_b, temperature=0.1, group=None)`` -- SimCLR NT-Xent over the embeddings of
  ALL ranks of ``group``.  The reference itself has no NT-Xent; this is
  the objective the north star specifies for that slot.
* ``byol_cosine_loss(preds, targets)`` -- the loss the reference actually trains with, fused fwd+bwd.

Both run only on CUDA through libmis_b200.so (tcgen05 kernels, csrc/ntxent.cu); there is no CPU or
PyTorch fallback.  The multi-GPU exchange is one all-gather of the normalised rows in the forward
and one all-gather of 2N log-sum-exp scalars in the backward (option L): no D-wide gradient
reduce-scatter is needed because every rank can form both P_ij and P_ji for its own rows.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib, peer
from ._lib import MIS_DTYPE_BF16, MIS_DTYPE_F32


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return MIS_DTYPE_F32
    if t.dtype == torch.bfloat16:
        return MIS_DTYPE_BF16
    raise TypeError(f"embeddings must be float32 or bfloat16, got {t.dtype}")


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _on_device:
    """`with torch.cuda.device(d)` costs ~10 us per entry; only switch when d is not already current."""

    def __init__(self, device):
        self.ctx = None if device.index is None or device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


class _LocalGraph:
    """prep -> forward -> backward of one shape captured into a CUDA graph over static buffers."""

    def __init__(self, z: torch.Tensor, inv_T: float, ws: torch.Tensor):
        self.z = torch.empty_like(z)
        self.loss = torch.empty((1,), dtype=torch.float32, device=z.device)
        self.dz = torch.empty_like(z)
        self.in_flight = 0
        self.z.copy_(z)
        CudaKernels._launch_local(self.z, inv_T, self.loss, self.dz, ws)          # warm-up outside the capture
        torch.cuda.current_stream(z.device).synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            CudaKernels._launch_local(self.z, inv_T, self.loss, self.dz, ws)

    def run(self, z: torch.Tensor):
        self.z.copy_(z)
        self.graph.replay()
        self.in_flight += 1
        return self.loss.clone(), self.dz, self


class _PeerGraph:
    """The multi-rank loss AND its gradient (grad_out = 1) captured into ONE CUDA graph over static buffers: prep with
    peer stores, forward tile + rows kernels, transpose, backward tile kernel(s) (which wait for the peers' log-sum-exp
    flags themselves) and the Jacobian kernel.  The first evaluation runs eagerly (a real, collective evaluation on every
    rank); the second is captured and replayed.  Epoch and buffer parity live in device memory, so a replay is a new
    evaluation.  A forward under ``torch.no_grad()`` has its own, forward-only graph."""

    def __init__(self, ex, z: torch.Tensor, inv_T: float):
        self.ex, self.inv_T = ex, float(inv_T)
        self.z = torch.empty_like(z)
        self.loss = torch.empty((1,), dtype=torch.float32, device=z.device)
        self.dz = torch.empty_like(z)
        self.ones = torch.ones((1,), dtype=torch.float32, device=z.device)
        self.graphs = {False: None, True: None}          # with_grad -> captured graph
        self.calls = {False: 0, True: 0}
        self.in_flight = 0                               # results whose backward() has not consumed self.dz yet

    def matches(self, z: torch.Tensor, inv_T: float) -> bool:
        return z.shape == self.z.shape and z.dtype == self.z.dtype and float(inv_T) == self.inv_T

    def _launch(self, with_grad: bool):
        CudaKernels._launch_fwd_peer(self.z, self.ex, self.inv_T, self.loss)
        if with_grad:
            CudaKernels._launch_bwd_peer(self.z, self.ex, self.inv_T, self.ones, self.dz)

    def run(self, z: torch.Tensor, with_grad: bool):
        self.z.copy_(z)
        if self.calls[with_grad] == 0:
            self._launch(with_grad)
        else:
            if self.graphs[with_grad] is None:
                torch.cuda.current_stream(z.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch(with_grad)
                self.graphs[with_grad] = g
            self.graphs[with_grad].replay()
        self.calls[with_grad] += 1
        if with_grad:
            self.in_flight += 1
        return self.loss.clone(), (self.dz if with_grad else None), self


class CudaKernels:
    """The three device steps of NT-Xent, bound to the C ABI (include/mis_b200.h)."""

    launches = 0   # kernels launched through this class (bench.py's gpu_launches claim)

    @staticmethod
    def launch_mode(world: int) -> str:
        """How the loss's kernels reach the GPU (reported by bench.py)."""
        if os.environ.get("MIS_NTXENT_GRAPH") == "1":
            if world == 1:
                return ("graph: one CUDA graph replay per step (prep, forward tile + rows kernels, transpose, backward tile + "
                        "Jacobian kernels)")
            return ("graph: one CUDA graph replay per step (prep with peer stores, forward tile + rows kernels, transpose, "
                    "backward tile + Jacobian kernels; the peers' flags are awaited inside the tile kernels; epoch and buffer "
                    "parity are read from device memory)")
        return "eager"

    @staticmethod
    def prep(z: torch.Tensor):
        if not z.is_cuda:
            raise RuntimeError("nt_xent_loss has no CPU path: embeddings must be CUDA tensors")
        z = z.contiguous()
        rows, D = z.shape
        rp = CudaKernels.padded_rows(rows)          # the rank's block of the gathered matrix: rows padded to 128
        u = torch.empty((rp, D), dtype=torch.float32, device=z.device)
        rinv = torch.empty((rp,), dtype=torch.float32, device=z.device)
        with _on_device(z.device):
            rc = _lib.lib.mis_ntxent_prep(z.data_ptr(), _dt(z), rows, D, u.data_ptr(), rinv.data_ptr(), _stream(z))
        _lib.check(rc, "mis_ntxent_prep")
        CudaKernels.launches += 1
        return z, u, rinv

    @staticmethod
    def padded_rows(rows: int) -> int:
        """Rows of one rank's block in the gathered matrix (mis_ntxent_padded_rows: the next multiple of 128)."""
        return int(_lib.lib.mis_ntxent_padded_rows(int(rows)))

    _workspaces: dict = {}   # (device, rows, D, stream) -> uint8 workspace of the single-rank fused path (stream-ordered reuse)
    _graphs: dict = {}       # (device, rows, D, dtype, inv_T, stream) -> captured single-rank step (MIS_NTXENT_GRAPH=1)

    @staticmethod
    def _workspace(device, rows: int, D: int) -> torch.Tensor:
        # one workspace per stream: reuse is stream-ordered, two streams must not share one
        key = (device, rows, D, torch.cuda.current_stream(device).cuda_stream)
        ws = CudaKernels._workspaces.get(key)
        if ws is None:
            n = int(_lib.lib.mis_ntxent_fwd_bwd_workspace_bytes(rows, D))
            if n <= 0:
                raise ValueError(f"nt_xent_loss: unsupported shape {(rows, D)}")
            ws = CudaKernels._workspaces[key] = torch.empty((n,), dtype=torch.uint8, device=device)
        return ws

    @staticmethod
    def _launch_local(z, inv_T, loss, dz, ws):
        rows, D = z.shape
        with _on_device(z.device):
            rc = _lib.lib.mis_ntxent_fwd_bwd(z.data_ptr(), _dt(z), rows, D, inv_T, loss.data_ptr(), dz.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _stream(z))
        _lib.check(rc, "mis_ntxent_fwd_bwd")

    @staticmethod
    def fwd_bwd_local(z: torch.Tensor, inv_T: float):
        """Single-rank loss and dL/dz in one call (mis_ntxent_fwd_bwd): 6 kernels + a memset, one host round trip.

        With ``MIS_NTXENT_GRAPH=1`` the seven launches are captured once per (shape, dtype, T) into a CUDA graph over
        static buffers and replayed (the step is launch-bound at SSL batch sizes: 6 kernels of 4-12 us).  The graph is
        bypassed while a previous result of the same shape still waits for its backward."""
        if not z.is_cuda:
            raise RuntimeError("nt_xent_loss has no CPU path: embeddings must be CUDA tensors")
        z = z.detach().contiguous()
        rows, D = z.shape
        ws = CudaKernels._workspace(z.device, rows, D)
        if os.environ.get("MIS_NTXENT_GRAPH", "0") == "1":
            key = (z.device, rows, D, z.dtype, float(inv_T), torch.cuda.current_stream(z.device).cuda_stream)
            g = CudaKernels._graphs.get(key)
            if g is None:
                g = CudaKernels._graphs[key] = _LocalGraph(z, inv_T, ws)
            if g.in_flight == 0:
                CudaKernels.launches += CudaKernels._n_fwd_bwd(D)
                return g.run(z)
        loss = torch.empty((1,), dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z)
        CudaKernels._launch_local(z, inv_T, loss, dz, ws)
        CudaKernels.launches += CudaKernels._n_fwd_bwd(D)
        return loss, dz, None

    @staticmethod
    def _n_bwd(D: int) -> int:
        """Kernels of one backward: transpose + tile kernel + Jacobian kernel (D <= 256: dU stays in TMEM), or transpose +
        tile kernel writing W + the dU GEMM + Jacobian kernel (wider embeddings)."""
        return 3 if D <= 256 else 4

    @staticmethod
    def _n_fwd_bwd(D: int) -> int:
        return 3 + CudaKernels._n_bwd(D)         # prep + forward tile kernel + forward rows kernel + backward

    @staticmethod
    def _peer_buffers(ex, z):
        if ex.scratch is None:
            rows, D = z.shape
            ex.scratch = CudaKernels.scratch(rows, ex.cols, D, z.device)
            ex.rinv = torch.empty((CudaKernels.padded_rows(rows),), dtype=torch.float32, device=z.device)

    @staticmethod
    def _launch_fwd_peer(z, ex, inv_T, loss):
        rows, D = z.shape
        with _on_device(z.device):
            rc = _lib.lib.mis_ntxent_fwd_peer(z.data_ptr(), _dt(z), rows, D, inv_T, ex.world, ex.rank, ex.u_peers[0],
                                              ex.u_peers[1], ex.l_peers[0], ex.l_peers[1], ex.ctl_peers, peer.timeout_s(),
                                              ex.rinv.data_ptr(), loss.data_ptr(), ex.scratch.data_ptr(),
                                              ex.scratch.numel(), _stream(z))
        _lib.check(rc, "mis_ntxent_fwd_peer")

    @staticmethod
    def _launch_bwd_peer(z, ex, inv_T, g, dz):
        rows, D = z.shape
        with _on_device(z.device):
            rc = _lib.lib.mis_ntxent_bwd_peer(z.data_ptr(), _dt(z), ex.rinv.data_ptr(), rows, D, inv_T, 1.0, g.data_ptr(),
                                              dz.data_ptr(), ex.world, ex.rank, ex.u_peers[0], ex.u_peers[1], ex.l_peers[0],
                                              ex.l_peers[1], ex.ctl_peers, peer.timeout_s(), ex.scratch.data_ptr(),
                                              ex.scratch.numel(), _stream(z))
        _lib.check(rc, "mis_ntxent_bwd_peer")

    @staticmethod
    def eval_peer(z: torch.Tensor, ex, inv_T: float, with_grad: bool):
        """Loss (and, with ``with_grad``, dL/dz for grad_out = 1) with the NVLink exchange fused into the kernels:
        mis_ntxent_fwd_peer (prep kernel with peer stores + ONE tile kernel that waits for the peers' rows itself + rows
        kernel) then mis_ntxent_bwd_peer (transpose + tile kernel(s) that wait for the peers' log-sum-exp flags + Jacobian
        kernel), issued back to back on the caller's stream -- every rank issues backward(k) before forward(k+1), which
        is what makes the two-deep buffer parity sufficient.  With ``MIS_NTXENT_GRAPH=1`` the whole evaluation is ONE
        CUDA graph replay over static buffers.  Returns (loss, dz or None, graph holder or None)."""
        if not z.is_cuda:
            raise RuntimeError("nt_xent_loss has no CPU path: embeddings must be CUDA tensors")
        z = z.detach().contiguous()
        CudaKernels._peer_buffers(ex, z)
        CudaKernels.launches += 3 + (CudaKernels._n_bwd(z.shape[1]) if with_grad else 0)
        if os.environ.get("MIS_NTXENT_GRAPH", "0") == "1":
            if ex.graph is None:
                ex.graph = _PeerGraph(ex, z, inv_T)
            if ex.graph.matches(z, inv_T) and ex.graph.in_flight == 0:
                return ex.graph.run(z, with_grad)
        loss = torch.empty((1,), dtype=torch.float32, device=z.device)
        CudaKernels._launch_fwd_peer(z, ex, inv_T, loss)
        dz = None
        if with_grad:
            dz = torch.empty_like(z)
            CudaKernels._launch_bwd_peer(z, ex, inv_T, CudaKernels._ones(z.device), dz)
        return loss, dz, None

    _ones_cache: dict = {}

    @staticmethod
    def _ones(device) -> torch.Tensor:
        t = CudaKernels._ones_cache.get(device)
        if t is None:
            t = CudaKernels._ones_cache[device] = torch.ones((1,), dtype=torch.float32, device=device)
        return t

    @staticmethod
    def scratch(rows: int, cols: int, D: int, device) -> torch.Tensor:
        n = int(_lib.lib.mis_ntxent_scratch_bytes(rows, cols, D))
        return torch.empty((n,), dtype=torch.uint8, device=device)

    @staticmethod
    def fwd(u_all: torch.Tensor, row0: int, rows: int, inv_T: float, scratch: torch.Tensor):
        cols, D = u_all.shape
        lse = torch.empty((CudaKernels.padded_rows(rows),), dtype=torch.float32, device=u_all.device)
        loss = torch.empty((1,), dtype=torch.float32, device=u_all.device)
        with _on_device(u_all.device):
            rc = _lib.lib.mis_ntxent_fwd(u_all.data_ptr(), cols, D, row0, rows, inv_T, lse.data_ptr(), loss.data_ptr(),
                                         scratch.data_ptr(), scratch.numel(), _stream(u_all))
        _lib.check(rc, "mis_ntxent_fwd")
        CudaKernels.launches += 2
        return lse, loss

    @staticmethod
    def bwd(u_all, lse_all, z, rinv, row0: int, inv_T: float, grad_out: torch.Tensor, scratch: torch.Tensor):
        cols, D = u_all.shape
        rows = z.shape[0]
        dz = torch.empty_like(z)
        g = grad_out.to(torch.float32).reshape(1).contiguous()
        with _on_device(z.device):
            rc = _lib.lib.mis_ntxent_bwd(u_all.data_ptr(), lse_all.data_ptr(), z.data_ptr(), _dt(z), rinv.data_ptr(),
                                         cols, D, row0, rows, inv_T, 1.0, g.data_ptr(), dz.data_ptr(),
                                         scratch.data_ptr(), scratch.numel(), _stream(z))
        _lib.check(rc, "mis_ntxent_bwd")
        CudaKernels.launches += CudaKernels._n_bwd(D)
        return dz


def _all_gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    """Rank-major concatenation along dim 0 (the layout of concat_all_gather, train/callback/knn.py:143-144)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


class _NTXent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, temperature: float, group, kernels):
        if z.dim() != 2 or z.shape[0] % 2:
            raise ValueError(f"expected [2*B_local, D] rows laid out [view1; view2], got {tuple(z.shape)}")
        distributed = group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        world = dist.get_world_size(group) if distributed else 1
        rank = dist.get_rank(group) if distributed else 0
        rows = z.shape[0]
        inv_T = 1.0 / float(temperature)
        if not distributed and kernels is CudaKernels and ctx.needs_input_grad[0]:
            # single rank: prep + forward + backward (grad_out = 1) in one ABI call; backward() only scales
            loss, dz, holder = kernels.fwd_bwd_local(z, inv_T)
            ctx.save_for_backward(dz)
            ctx.meta = None
            ctx.holder = holder                          # the captured graph whose static dz this result refers to
            return loss.reshape(())
        if distributed and kernels is CudaKernels:
            ex = peer.get_exchange(group, rows, z.shape[1], z.device)
            if ex is not None:
                # forward AND backward kernels are issued here (the gradient for grad_out = 1); backward() only scales.
                # So no evaluation is ever left half-way through the exchange, whatever the caller does with the loss.
                loss, dz, holder = kernels.eval_peer(z, ex, inv_T, bool(ctx.needs_input_grad[0]))
                if dz is not None:
                    ctx.save_for_backward(dz)
                ctx.meta = None
                ctx.holder = holder
                return loss.reshape(())
        z, u, rinv = kernels.prep(z)
        u_all = _all_gather_rows(u, group) if distributed else u
        scratch = kernels.scratch(rows, u_all.shape[0], u_all.shape[1], z.device)
        lse, loss = kernels.fwd(u_all, rank * kernels.padded_rows(rows), rows, inv_T, scratch)
        ctx.save_for_backward(z, u_all, rinv, lse)
        ctx.meta = (inv_T, group, distributed, rank, kernels, scratch)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.meta is None:
            (dz,) = ctx.saved_tensors
            out = dz * grad_out.to(dz.dtype)
            if ctx.holder is not None:                   # the captured graph's static dz is free again (once per result)
                ctx.holder.in_flight -= 1
                ctx.holder = None
            return out, None, None, None
        z, u_all, rinv, lse = ctx.saved_tensors
        inv_T, group, distributed, rank, kernels, scratch = ctx.meta
        lse_all = _all_gather_rows(lse, group) if distributed else lse
        dz = kernels.bwd(u_all, lse_all, z, rinv, rank * kernels.padded_rows(z.shape[0]), inv_T, grad_out, scratch)
        return dz, None, None, None


def nt_xent_rows(z: torch.Tensor, temperature: float = 0.1, group=None, _kernels=CudaKernels) -> torch.Tensor:
    """NT-Xent for rows already laid out ``[view1_local; view2_local]`` (the output of the encoder on
    ``cat([view1, view2])``, byol_pytorch.py:207-208).  Returns this rank's loss L_r; its
    backward yields sum_r' dL_r'/dz_local so that DDP's 1/W gradient averaging is exact."""
    return _NTXent.apply(z, float(temperature), group, _kernels)


def nt_xent_loss(z_a: torch.Tensor, z_b: torch.Tensor, temperature: float = 0.1, group=None) -> torch.Tensor:
    """Drop-in for the ``loss = self.cosine_similarity_loss(preds, targets)`` slot (byol_pytorch.py:217):
    z_a[i] and z_b[i] are the two views of image i."""
    if z_a.shape != z_b.shape:
        raise ValueError(f"shape mismatch {tuple(z_a.shape)} vs {tuple(z_b.shape)}")
    if group is None and dist.is_available() and dist.is_initialized():
        group = dist.group.WORLD
    return nt_xent_rows(torch.cat([z_a, z_b], dim=0), temperature, group)


class _BYOLCosine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, targets):
        if not preds.is_cuda:
            raise RuntimeError("byol_cosine_loss has no CPU path: inputs must be CUDA tensors")
        if preds.shape != targets.shape or preds.dim() != 2:
            raise ValueError(f"expected two [rows, D] tensors, got {tuple(preds.shape)} and {tuple(targets.shape)}")
        p = preds.detach().to(torch.float32).contiguous()
        t = targets.detach().to(torch.float32).contiguous()
        rows, D = p.shape
        loss = torch.empty((1,), dtype=torch.float32, device=p.device)
        dp = torch.empty_like(p)
        scratch = torch.empty((rows,), dtype=torch.float32, device=p.device)
        with _on_device(p.device):
            rc = _lib.lib.mis_byol_loss_fwd_bwd(p.data_ptr(), t.data_ptr(), rows, D, loss.data_ptr(), dp.data_ptr(),
                                                scratch.data_ptr(), _stream(p))
        _lib.check(rc, "mis_byol_loss_fwd_bwd")
        ctx.save_for_backward(dp)
        ctx.in_dtype = preds.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (dp,) = ctx.saved_tensors
        return (dp * grad_out).to(ctx.in_dtype), None


def byol_cosine_loss(preds: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """``BYOL.cosine_similarity_loss`` (byol_pytorch.py:181-198): 2 - 2*mean cos(preds_i, targets_i).
    Gradient flows to ``preds`` only (the reference computes targets under no_grad, :212-214)."""
    return _BYOLCosine.apply(preds, targets)
