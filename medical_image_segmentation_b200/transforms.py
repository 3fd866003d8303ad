"""Drop-in for the reference's two-view transform object, running as one fused CUDA kernel.

Mirrors ``BYOLRGBDataTransforms`` (train/data_loaders/lightning_module.py:39-64): same constructor
arguments, ``__call__(x) -> [view1, view2]``.  Differences, all forced by moving the chain from a
per-sample CPU callable to a per-batch GPU kernel:

* ``x`` is a *batch* ``[B, C, H, W]`` (or ``[B, H, W]``), C = 1 (16-bit slices, the fused single-kernel path) or C = 3
  (saturation / hue / RandomGrayscale live; a resample kernel plus a colour kernel, crop a multiple of 8 up to 256), of
  raw ``torch.uint16`` slices -- on the GPU, or
  on the host (then it is copied through pinned memory first).  The reference feeds one decoded
  image at a time to DataLoader workers.
* the two views come back as ``[B, C, s, s]`` tensors that are the two halves of ONE
  ``[2B, C, s, s]`` buffer (``.views_buffer``), so the ``torch.cat(views)`` of
  train/model/byol_pytorch.py:207 is free.
* random parameters are drawn from torch's global CPU generator in exactly the reference's order
  (params.py), so ``torch.manual_seed(k)`` gives the same crops / flips / jitter as the reference.
* ``use_tma`` selects the K1 variant: ``False``/``0`` (default: the strip kernel, csrc/aug_strip.cu, wherever it
  applies), ``True``/``1`` (band kernel, 2-D TMA tensor-map boxes fed by a producer warp), ``2`` (band kernel,
  per-thread cp.async rings, 16-row sub-bands) or ``3`` (round-1 warp-tile kernel, csrc/aug_tile.cu).  All four are
  parity-tested against the same oracle.
* ``prefetch_params=True`` draws the next batch's parameters on a helper thread while the GPU works on the current
  one (same records in the same order; see ``next_params``).
* ``blur_prob`` / ``solarize_prob`` default to the reference's ``(1.0, 0.1)`` / ``(0.0, 0.2)``
  (lightning_module.py:40).  GaussianBlur(23) runs as a second kernel over the views that drew it
  (``mis_aug_blur_views``), RandomSolarize in K1's store epilogue; both need the strip kernel (``use_tma=0``, one
  channel).  ``RandomSolarize(128)`` is written for 0..255 images: on the [0,1] scale the threshold is 128/255
  (torchvision itself refuses 128 on a float image, functional/_color.py:498-499; the reference only ever feeds it
  uint8 PIL images).  ``RandomGrayscale`` is the identity for one channel.
* ``generator``: a private ``torch.Generator`` to draw the parameters from instead of torch's global CPU generator
  (recommended with ``prefetch_params=True``: the helper thread then cannot be disturbed by, or disturb, other users
  of the global stream).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MIS_DTYPE_BF16, MIS_DTYPE_F32, MIS_VIEW_BLUR, MIS_VIEW_SOLARIZE, VIEW_PARAMS_DTYPE
from .params import draw_two_view_params


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

_U16_MAX = 65535.0


def _as_float_seq(v, n: int, name: str) -> list[float]:
    if isinstance(v, (int, float)):
        v = [float(v)] * n
    v = [float(t) for t in v]
    if len(v) != n:
        raise ValueError(f"{name} has {len(v)} entries for {n} channel(s)")
    return v


class FusedTwoViewTransforms:
    def __init__(self, crop_size: int, mean: Sequence[float], std: Sequence[float],
                 blur_prob=(1.0, 0.1), solarize_prob=(0.0, 0.2), *, out_dtype=torch.bfloat16,
                 window: tuple[float, float] | None = None, use_tma: bool | int = False,
                 prefetch_params: bool = False, generator: torch.Generator | None = None):
        assert len(blur_prob) == 2 and len(solarize_prob) == 2, "atm only 2 views are supported"
        if any(not 0.0 <= float(p) <= 1.0 for p in tuple(blur_prob) + tuple(solarize_prob)):
            raise ValueError("blur_prob / solarize_prob must be probabilities")
        if isinstance(crop_size, (tuple, list)):
            if len(crop_size) != 2 or crop_size[0] != crop_size[1]:
                raise NotImplementedError("only square crops are implemented")
            crop_size = crop_size[0]
        self.crop_size = int(crop_size)
        self.mean = mean
        self.std = std
        self.blur_prob = tuple(float(p) for p in blur_prob)
        self.solarize_prob = tuple(float(p) for p in solarize_prob)
        if out_dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("out_dtype must be torch.bfloat16 or torch.float32")
        self.out_dtype = out_dtype
        self.window = (0.0, _U16_MAX) if window is None else (float(window[0]), float(window[1]))
        self.use_tma = int(use_tma)
        if self.use_tma not in (0, 1, 2, 3):
            raise ValueError("use_tma must be 0 (strip kernel), 1 (TMA band kernel), 2 (cp.async band kernel) or "
                             "3 (warp-tile kernel)")
        self.prefetch_params = bool(prefetch_params)
        self.generator = generator
        self._pool = None                     # one helper thread drawing the NEXT batch's parameters
        self._pending = None                  # ((B, H, W), future)
        self._norm_cache = {}
        self._c_flags, self._c_bad, self._c_nl = C.c_uint32(0), C.c_int(-1), C.c_int(0)
        self._x_stage: torch.Tensor | None = None      # device staging buffer of host batches
        self._x_host_ring = []                          # (pinned host batch, event after its copies): alive while in flight
        self.last_h2d_bytes = 0
        self.views_buffer: torch.Tensor | None = None
        self.last_params: np.ndarray | None = None
        self.launches = 0
        self._staging = []        # ring of (pinned host table, device table, event): params stay alive while in flight
        self._staging_idx = 0

    # -- parameters ---------------------------------------------------------------------------
    def draw_params(self, B: int, H: int, W: int) -> np.ndarray:
        """Image-major records [2*i+v], drawn like B calls of the reference's __call__."""
        return draw_two_view_params(B, H, W, self.blur_prob, self.solarize_prob, generator=self.generator)

    def next_params(self, B: int, H: int, W: int, view_major: bool = False) -> np.ndarray:
        """``draw_params`` with the draw of the FOLLOWING batch started on a helper thread (``prefetch_params=True``):
        the host RNG replay (~0.4 ms per 1024 slices) then overlaps the GPU work of the current batch.  The records are
        the same, in the same order, as back-to-back ``draw_params`` calls -- as long as nothing else consumes torch's
        global CPU generator between calls (the helper thread reads and writes its state); pass ``generator=`` to the
        constructor to draw from a private generator instead.  ``view_major=True`` returns ``to_view_major(...)`` of
        the table (reordered on the helper thread as well)."""
        if not self.prefetch_params:
            return self._draw(B, H, W, view_major)
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mis-params")
        key = (B, H, W, bool(view_major))
        params = None
        if self._pending is not None:
            pkey, fut = self._pending
            self._pending = None
            got = fut.result()                       # always joined: the generator is never touched concurrently
            if pkey == key:
                params = got
            elif pkey[:3] == key[:3]:                # same draw, other order: reorder instead of drawing again
                params = self.to_view_major(got) if view_major else self._to_image_major(got)
        if params is None:
            params = self._draw(B, H, W, view_major)
        self._pending = (key, self._pool.submit(self._draw, B, H, W, view_major))
        return params

    def _draw(self, B: int, H: int, W: int, view_major: bool) -> np.ndarray:
        p = self.draw_params(B, H, W)
        return self.to_view_major(p) if view_major else p

    @staticmethod
    def _to_image_major(vm: np.ndarray) -> np.ndarray:
        B = vm.shape[0] // 2
        out = np.empty_like(vm)
        out[0::2], out[1::2] = vm[:B], vm[B:]
        return out

    def drain_prefetch(self) -> None:
        """Wait for (and drop) a parameter draw in flight, e.g. before re-seeding the generator."""
        if self._pending is not None:
            self._pending[1].result()
            self._pending = None

    @staticmethod
    def to_view_major(params: np.ndarray) -> np.ndarray:
        """[2*i+v] -> [v*B+i]: row order of cat([view1, view2]) (byol_pytorch.py:207)."""
        B = params.shape[0] // 2
        src = np.ascontiguousarray(params)
        out = np.empty_like(src)
        _lib.check(_lib.lib.mis_params_to_view_major(src.ctypes.data, B, out.ctypes.data), "mis_params_to_view_major")
        return out

    # -- device path --------------------------------------------------------------------------
    def apply(self, x: torch.Tensor, params_view_major: np.ndarray, out: torch.Tensor | None = None) -> torch.Tensor:
        """Run K1 for an explicit parameter table.  Returns the [n_views, C, s, s] buffer."""
        if x.dtype != torch.uint16:
            raise TypeError(f"expected raw torch.uint16 slices, got {x.dtype}")
        if not x.is_cuda:
            raise ValueError("apply() needs a CUDA tensor (use __call__ for host batches)")
        if x.dim() == 3:
            x = x[:, None]
        if x.dim() != 4:
            raise ValueError(f"expected [B,C,H,W], got {tuple(x.shape)}")
        x = x.contiguous()
        B, Cc, H, W = x.shape
        mean_c, std_c = self._norm_args(Cc)
        n_views = int(params_view_major.shape[0])
        assert params_view_major.dtype == VIEW_PARAMS_DTYPE
        s = self.crop_size
        if out is None:
            out = torch.empty((n_views, Cc, s, s), dtype=self.out_dtype, device=x.device)
        else:
            assert out.is_cuda and out.is_contiguous() and out.dtype == self.out_dtype
            assert tuple(out.shape) == (n_views, Cc, s, s)
        if n_views == 0:
            return out
        p = params_view_major if params_view_major.flags.c_contiguous else np.ascontiguousarray(params_view_major)
        # check + staging copy + launch order + H2D + K1 (+ blur kernel) in one native call (mis_aug_two_view_staged)
        host, dev, ev = self._staging_slot(p.nbytes + 4 * n_views, x.device)
        flags, bad, nl = self._c_flags, self._c_bad, self._c_nl
        with _on_device(x.device):
            stream = torch.cuda.current_stream(x.device)
            rc = _lib.lib.mis_aug_two_view_staged(
                x.data_ptr(), B, Cc, H, W, Cc * H * W, p.ctypes.data, n_views, host.data_ptr(), dev.data_ptr(),
                host.numel(), self.window[0], self.window[1], mean_c, std_c, out.data_ptr(), s,
                MIS_DTYPE_F32 if self.out_dtype == torch.float32 else MIS_DTYPE_BF16, self.use_tma,
                C.byref(flags), C.byref(bad), C.byref(nl), C.c_void_p(stream.cuda_stream))
            ev.record(stream)
        _lib.check(rc, "mis_aug_two_view_staged")
        if bad.value >= 0:
            self._validate_table(p, B, H, W)             # raises the ValueError that names the record
        self.launches += nl.value
        return out

    def _staging_slot(self, nbytes: int, device):
        """One of three persistent (pinned, device) block pairs for the table and its launch order: no per-call
        cudaHostAlloc; a block is reused only after the copy that last read it has completed."""
        if len(self._staging) < 3:
            slot = [None, None, torch.cuda.Event()]
            self._staging.append(slot)
        else:
            slot = self._staging[self._staging_idx % 3]
            self._staging_idx += 1
            slot[2].synchronize()
        if slot[0] is None or slot[0].numel() < nbytes or slot[1].device != torch.device(device):
            slot[0] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            slot[1] = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return slot

    def _norm_args(self, Cc: int):
        """mean / std as C float arrays (cached per channel count; the arrays stay alive with the transform)."""
        got = self._norm_cache.get(Cc)
        if got is None:
            mean = (C.c_float * Cc)(*_as_float_seq(self.mean, Cc, "mean"))
            std = (C.c_float * Cc)(*_as_float_seq(self.std, Cc, "std"))
            got = self._norm_cache[Cc] = (mean, std, C.cast(mean, C.c_void_p), C.cast(std, C.c_void_p))
        return got[2], got[3]

    @staticmethod
    def _validate_table(p: np.ndarray, B: int, H: int, W: int) -> int:
        """The kernel trusts the table: check every record addresses a slice of the batch and a box inside it
        (mis_view_params_check, a native pass of a few microseconds).  Returns the OR of all flag words."""
        if p.shape[0] == 0:
            return 0
        p = np.ascontiguousarray(p)
        flags, bad = C.c_uint32(0), C.c_int(-1)
        _lib.check(_lib.lib.mis_view_params_check(p.ctypes.data, p.shape[0], B, H, W, C.byref(flags), C.byref(bad)),
                   "mis_view_params_check")
        if bad.value >= 0:
            r = p[bad.value]
            if (r["flags"] & MIS_VIEW_BLUR) and not (np.isfinite(r["blur_sigma"]) and r["blur_sigma"] > 0):
                raise ValueError(f"view record {bad.value}: a record with MIS_VIEW_BLUR needs a positive blur_sigma")
            raise ValueError(f"view record {bad.value} is outside the batch: img {int(r['img'])} of {B}, box (top "
                             f"{int(r['top'])}, left {int(r['left'])}, h {int(r['h'])}, w {int(r['w'])}) in {H}x{W}")
        return int(flags.value)

    def stage_needed_rows(self, x_host: torch.Tensor, params: np.ndarray, device=None,
                          min_gap_bytes: int = 256 << 10, out: torch.Tensor | None = None) -> torch.Tensor:
        """Host batch -> device, copying only the rows the records read (mis_h2d_needed_rows); ranges closer than
        ``min_gap_bytes`` are merged into one copy (a copy costs ~200 KB of PCIe time to set up).  Returns the full-size
        device buffer K1 reads; rows no record touches are stale.  ``self.last_h2d_bytes`` = bytes put on the wire.
        The copies go to torch's current stream; ``out`` names the device buffer (e.g. one of two, filled on a copy
        stream while the previous batch is being processed), default: one buffer owned by the transform."""
        if x_host.dtype != torch.uint16 or x_host.is_cuda or x_host.dim() != 4:
            raise TypeError("stage_needed_rows expects a host torch.uint16 [B,C,H,W] batch")
        x_host = x_host.contiguous()
        if not x_host.is_pinned():
            x_host = x_host.pin_memory()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        B, Cc, H, W = x_host.shape
        if out is not None:
            if out.dtype != torch.uint16 or out.shape != x_host.shape or out.device != device or not out.is_contiguous():
                raise ValueError("stage_needed_rows: `out` must be a contiguous torch.uint16 device tensor of the batch's shape")
            stage = out
        else:
            if self._x_stage is None or self._x_stage.shape != x_host.shape or self._x_stage.device != device:
                self._x_stage = torch.empty(x_host.shape, dtype=torch.uint16, device=device)
            stage = self._x_stage
        p = np.ascontiguousarray(params)
        nbytes = C.c_int64(0)
        with _on_device(device):
            rc = _lib.lib.mis_h2d_needed_rows(x_host.data_ptr(), stage.data_ptr(), B, Cc, H, W, Cc * H * W,
                                              p.ctypes.data, p.shape[0], int(min_gap_bytes),
                                              C.c_void_p(torch.cuda.current_stream(device).cuda_stream), C.byref(nbytes))
        _lib.check(rc, "mis_h2d_needed_rows")
        # The raw cudaMemcpyAsync calls inside the library are invisible to torch's pinned-memory allocator: every host
        # batch stays referenced until an event recorded behind its copies has completed (a lagging GPU may still have
        # copies of older batches queued), so a freshly pinned block can never be recycled under a pending copy.
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self._x_host_ring = [(h, e) for (h, e) in self._x_host_ring if not e.query()]
        self._x_host_ring.append((x_host, ev))
        self.last_h2d_bytes = int(nbytes.value)
        return stage

    def __call__(self, x) -> list[torch.Tensor]:
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        if x.dim() == 2:
            x = x[None, None]
        elif x.dim() == 3:
            x = x[:, None]
        if not x.is_cuda and not torch.cuda.is_available():
            raise RuntimeError("FusedTwoViewTransforms has no CPU path: a CUDA device is required")
        B, _, H, W = x.shape
        params = self.next_params(B, H, W)
        self.last_params = params
        if not x.is_cuda:
            x = self.stage_needed_rows(x, params)       # host batch: only the rows the crops read cross PCIe
        out = self.apply(x, self.to_view_major(params))
        self.views_buffer = out
        return [out[:B], out[B:]]


class FusedFFCVTwoViewTransforms(FusedTwoViewTransforms):
    """The FFCV flavour of the chain: ``BYOLRGBFFCVDataTransforms(device, crop_size, mean, std, solarize_prob)``
    (train/data_loaders/lightning_module.py:67-98), whose per-view pipeline is RandomResizedCrop(scale (0.08, 1), ratio
    (3/4, 4/3)) -> RandomHorizontalFlip(0.5) -> RandomGrayscale(0.2) -> RandomSolarization(p, 128) -> Normalize --
    colour jitter and GaussianBlur are commented out there (:81-86).

    The reference runs those ops as FFCV numba stages on the CPU, followed by an H2D copy; FFCV-SSL is an un-vendored
    fork that is not installed here, so its pipelines cannot be built.  What can be made real without it is the other end
    of the seam: the loader yields the RAW uint16 batch (an ``NDArrayField`` instead of the 8-bit ``RGBImageField``),
    Lightning moves it to the GPU, and ``on_after_batch_transfer`` produces ``(view_1, labels, view_2)`` -- the tuple
    ``BYOL.training_step`` unpacks (byol_pytorch.py:201-204) -- with the fused kernel, the same op set and the same
    probabilities.  Two deliberate differences: the parameters come from torch's CPU generator (FFCV draws from numpy's
    inside numba; that stream cannot be reproduced without its source), and the resample is torchvision's antialiased
    bilinear filter rather than FFCV's ``cv::resize`` (recalled, not verifiable here: SURVEY A.6).

    ``get_transforms()`` keeps the reference's return shape -- one pipeline (list of stages) per view; a stage is a
    callable on the raw device batch.  Both pipelines share one kernel launch per batch.
    """

    def __init__(self, device, crop_size, mean, std, solarize_prob=(0.0, 0.2), *, out_dtype=torch.bfloat16, **kw):
        if isinstance(crop_size, (tuple, list)) and len(crop_size) == 2 and crop_size[0] == crop_size[1]:
            crop_size = crop_size[0]
        super().__init__(crop_size, mean, std, blur_prob=(0.0, 0.0), solarize_prob=solarize_prob, out_dtype=out_dtype, **kw)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._cached = None            # (batch identity, [view1, view2])

    def draw_params(self, B: int, H: int, W: int) -> np.ndarray:
        """Crop box, flip, grayscale and solarize draws of the FFCV op set: the jitter flag is cleared (the op is not
        in this pipeline) -- torch's stream is used, so the draw ORDER is the torchvision chain's."""
        p = super().draw_params(B, H, W)
        p["flags"] &= ~np.uint32(2 | 8)
        p["order"] = (0, 1, 2, 3)
        p["brightness"] = p["contrast"] = p["saturation"] = 1.0
        p["hue"] = 0.0
        return p

    class _Stage:
        def __init__(self, owner, view: int):
            self.owner, self.view = owner, view

        def __call__(self, x):
            return self.owner._views_of(x)[self.view]

    def _views_of(self, x):
        key = (x.data_ptr(), tuple(x.shape), x._version)
        if self._cached is None or self._cached[0] != key:
            self._cached = (key, FusedTwoViewTransforms.__call__(self, x))
        return self._cached[1]

    def get_transforms(self):
        return [[self._Stage(self, 0)], [self._Stage(self, 1)]]

    def on_after_batch_transfer(self, batch, dataloader_idx: int = 0):
        """Lightning hook: ``batch = (x_u16 [B,C,H,W] on the GPU, labels, ...)`` -> ``(view_1, labels, view_2)``."""
        x, labels = batch[0], batch[1]
        if not x.is_cuda:
            x = x.to(self.device, non_blocking=True)
        v1, v2 = FusedTwoViewTransforms.__call__(self, x)
        return v1, labels, v2


def algorithmic_bytes(params: np.ndarray, C_: int, crop_size: int, out_dtype=torch.bfloat16) -> int:
    """Bytes K1 must move for this table: u16 crop windows read once + output written once."""
    p = np.ascontiguousarray(params)
    return int(_lib.lib.mis_aug_algorithmic_bytes(p.ctypes.data, p.shape[0], C_, crop_size,
                                                  MIS_DTYPE_F32 if out_dtype == torch.float32 else MIS_DTYPE_BF16))


class FusedResizeJitterTransforms(FusedTwoViewTransforms):
    """Single-view flavour of the chain used by the reference's Decathlon data module
    (train/data_loaders/lightning_module.py:684-712): ``Resize((s, s))`` (antialiased bilinear) ->
    ``ColorJitter(brightness, contrast)`` (train) or nothing (default/val) -> ``ToDtype(float32, scale=True)`` ->
    ``Normalize(mean, std)``.  Runs on the same fused kernel K1 with one record per image (box = whole slice).

    ``__call__(x) -> [B, C, s, s]``.  ``brightness=contrast=None`` reproduces ``default_transforms`` (no RNG use).
    """

    def __init__(self, size, mean, std, brightness=None, contrast=None, *, out_dtype=torch.bfloat16, window=None,
                 use_tma: bool | int = False):
        super().__init__(size, mean, std, (0.0, 0.0), (0.0, 0.0), out_dtype=out_dtype, window=window, use_tma=use_tma)
        self.brightness = brightness
        self.contrast = contrast

    def draw_params(self, B: int, H: int, W: int) -> np.ndarray:
        out = np.zeros(B, VIEW_PARAMS_DTYPE)
        if self.brightness is None and self.contrast is None:          # default_transforms: no ColorJitter in the chain
            out["img"] = np.arange(B)
            out["h"], out["w"] = H, W
            out["order"] = (0, 1, 2, 3)
            out["brightness"] = out["contrast"] = out["saturation"] = 1.0
            return out
        state = torch.get_rng_state()
        blob = state.numpy()
        rc = _lib.lib.mis_draw_resize_jitter_params(blob.ctypes.data, blob.nbytes, B, 0, H, W,
                                                    float(self.brightness or 0.0), float(self.contrast or 0.0),
                                                    out.ctypes.data)
        _lib.check(rc, "mis_draw_resize_jitter_params")
        torch.set_rng_state(state)
        return out

    def __call__(self, x) -> torch.Tensor:
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        if x.dim() == 2:
            x = x[None, None]
        elif x.dim() == 3:
            x = x[:, None]
        if not x.is_cuda:
            if not torch.cuda.is_available():
                raise RuntimeError("FusedResizeJitterTransforms has no CPU path: a CUDA device is required")
            x = (x if x.is_pinned() else x.pin_memory()).cuda(non_blocking=True)
        B, _, H, W = x.shape
        params = self.draw_params(B, H, W)
        self.last_params = params
        out = self.apply(x, params)
        self.views_buffer = out
        return out
