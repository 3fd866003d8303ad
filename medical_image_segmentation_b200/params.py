"""Augmentation parameters of the two-view chain, drawn from torch's global CPU generator.

``draw_two_view_params`` consumes the generator exactly as ``n`` calls of the reference's
``BYOLRGBDataTransforms.__call__`` (train/data_loaders/lightning_module.py:63-64) would -- crop
boxes, flips, jitter flags, op order and factors are bit-identical for the same
``torch.manual_seed`` -- but does it natively (csrc/rng_replay.cu) at ~50 ns per draw instead of
one Python tensor op per draw.  ``draw_two_view_params_torch`` is the same replay written with
torch calls; it is used for the rare image the native code hands back (crop box depending on the
last bit of torch.exp) and as the cross-check in the tests.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import (MIS_VIEW_BLUR, MIS_VIEW_FLIP, MIS_VIEW_GRAY, MIS_VIEW_JITTER, MIS_VIEW_SOLARIZE,
                   VIEW_PARAMS_DTYPE)

# constants of the reference chain (lightning_module.py:44,49-52; torchvision RandomResizedCrop defaults)
_RRC_SCALE = (0.08, 1.0)
_RRC_RATIO = (3.0 / 4.0, 4.0 / 3.0)
_JITTER = dict(brightness=(0.6, 1.4), contrast=(0.6, 1.4), saturation=(0.8, 1.2), hue=(-0.1, 0.1))


def _u(lo, hi, gen=None) -> float:
    return torch.empty(1).uniform_(lo, hi, generator=gen).item()


def _draw_view_torch(rec, H: int, W: int, blur_p: float, sol_p: float, gen=None) -> None:
    """One view, torch calls in torchvision's order (v2/_geometry.py:272-308, _transform.py:181,
    _container.py:104, _color.py:146-154)."""
    area = H * W
    log_ratio = torch.log(torch.tensor(_RRC_RATIO))
    for _ in range(10):
        target_area = area * _u(_RRC_SCALE[0], _RRC_SCALE[1], gen)
        aspect = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1], generator=gen)).item()
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= W and 0 < h <= H:
            top = torch.randint(0, H - h + 1, size=(1,), generator=gen).item()
            left = torch.randint(0, W - w + 1, size=(1,), generator=gen).item()
            break
    else:
        in_ratio = float(W) / float(H)
        if in_ratio < min(_RRC_RATIO):
            w = W
            h = int(round(w / min(_RRC_RATIO)))
        elif in_ratio > max(_RRC_RATIO):
            h = H
            w = int(round(h * max(_RRC_RATIO)))
        else:
            w, h = W, H
        top, left = (H - h) // 2, (W - w) // 2
    rec["top"], rec["left"], rec["h"], rec["w"] = top, left, h, w
    flags = 0
    rec["order"] = (0, 1, 2, 3)
    rec["brightness"], rec["contrast"], rec["saturation"], rec["hue"] = 1.0, 1.0, 1.0, 0.0
    if not bool(torch.rand(1, generator=gen) >= 0.5):
        flags |= MIS_VIEW_FLIP
    if not bool(torch.rand(1, generator=gen) >= 0.8):
        flags |= MIS_VIEW_JITTER
        rec["order"] = torch.randperm(4, generator=gen).numpy().astype(np.uint8)
        rec["brightness"] = _u(*_JITTER["brightness"], gen)
        rec["contrast"] = _u(*_JITTER["contrast"], gen)
        rec["saturation"] = _u(*_JITTER["saturation"], gen)
        rec["hue"] = _u(*_JITTER["hue"], gen)
    if not bool(torch.rand(1, generator=gen) >= 0.2):    # RandomGrayscale(p=0.2): identity for one channel
        flags |= MIS_VIEW_GRAY
    rec["blur_sigma"] = 0.0
    if not bool(torch.rand(1, generator=gen) >= blur_p):  # RandomApply([GaussianBlur(23)]); sigma: v2/_misc.py:209-211
        flags |= MIS_VIEW_BLUR
        rec["blur_sigma"] = _u(0.1, 2.0, gen)
    if not bool(torch.rand(1, generator=gen) >= sol_p):   # RandomSolarize(128)
        flags |= MIS_VIEW_SOLARIZE
    rec["flags"] = flags


def draw_two_view_params_torch(n_images: int, H: int, W: int, blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0),
                               img0: int = 0, generator: torch.Generator | None = None) -> np.ndarray:
    out = np.zeros(2 * n_images, VIEW_PARAMS_DTYPE)
    for i in range(n_images):
        for v in range(2):
            rec = out[2 * i + v]
            rec["img"] = img0 + i
            _draw_view_torch(rec, H, W, blur_prob[v], solarize_prob[v], generator)
    return out


@C.CFUNCTYPE(C.c_float, C.c_float, C.c_void_p)
def _torch_exp_f32(x, _ctx):
    """torch.exp on a one-element float32 tensor: the reference's own evaluation of the aspect ratio
    (torchvision v2/_geometry.py:284).  Called by the native replay for the rare draw whose crop box depends on its last
    bit (SLEEF's vector expf and libm's differ there)."""
    return float(torch.exp(torch.tensor([x], dtype=torch.float32))[0])


def draw_two_view_params(n_images: int, H: int, W: int, blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0),
                         out: np.ndarray | None = None, generator: torch.Generator | None = None) -> np.ndarray:
    """2*n_images records (image-major: [2*i+v]) drawn from torch's global CPU generator, or from ``generator`` (a
    private ``torch.Generator``: nothing else can then move the stream between two draws)."""
    get_state = torch.get_rng_state if generator is None else generator.get_state
    set_state = torch.set_rng_state if generator is None else generator.set_state
    if out is None:
        out = np.zeros(2 * n_images, VIEW_PARAMS_DTYPE)
    assert out.dtype == VIEW_PARAMS_DTYPE and out.shape == (2 * n_images,) and out.flags.c_contiguous
    bp = (C.c_float * 2)(*[float(p) for p in blur_prob])
    sp = (C.c_float * 2)(*[float(p) for p in solarize_prob])
    n_done = C.c_int(0)
    state = get_state()
    blob = state.numpy()
    done = 0
    while done < n_images:
        rc = _lib.lib.mis_draw_two_view_params_cb(blob.ctypes.data, blob.nbytes, n_images - done, done, H, W,
                                                  C.cast(bp, C.c_void_p), C.cast(sp, C.c_void_p),
                                                  out[2 * done:].ctypes.data, C.byref(n_done),
                                                  C.cast(_torch_exp_f32, C.c_void_p), None)
        _lib.check(rc, "mis_draw_two_view_params_cb")
        done += n_done.value
        if done < n_images:
            # (only without the callback) this image's crop box depends on the last bit of torch.exp: let torch draw it
            set_state(state)
            out[2 * done:2 * done + 2] = draw_two_view_params_torch(1, H, W, blur_prob, solarize_prob, img0=done,
                                                                    generator=generator)
            state = get_state()
            blob = state.numpy()
            done += 1
    set_state(state)
    return out
