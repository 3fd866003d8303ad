"""CPU oracle for the two-view augmentation path (TEST INFRASTRUCTURE ONLY).

Two independent restatements of what the reference computes for one image:

``TwoViewChainTV``
    The reference's transform *chain* (train/data_loaders/lightning_module.py:39-64)
    re-assembled from the same third-party library it uses (torchvision.transforms.v2).
    It consumes torch's global CPU generator exactly like the reference, so with the
    same ``torch.manual_seed`` it produces the same two views.  This is also what the
    CPU baseline times (it *is* the reference's CPU path: per-sample, per-op).

``draw_two_view_params`` + ``apply_view``
    A line-by-line numpy restatement of the arithmetic *inside* that chain:
      * the RNG draw order              (torchvision
        transforms/v2/_geometry.py:272-308, _transform.py:171-194,
        _container.py:101-110, _color.py:146-171)
      * antialiased bilinear resampling (A.2; ATen _upsample_bilinear2d_aa reached
        from transforms/v2/functional/_geometry.py:324-330)
      * brightness / contrast / normalise (A.3; functional/_color.py:92-97,114-125,
        190-205 and functional/_misc.py:37-67)
    This is the specification the CUDA kernel is written against: it exposes the crop
    box, flip flag, op order and factors as integers/floats (bit-exact gates) and the
    pixels as float32 (1e-3 gate).

Pinned against the reference itself by oracle/make_golden.py -> tests/golden/.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# ----------------------------------------------------------------------------------
# constants of the reference chain (lightning_module.py:44,50-51)
# ----------------------------------------------------------------------------------
RRC_SCALE = (0.08, 1.0)
RRC_RATIO = (3.0 / 4.0, 4.0 / 3.0)
FLIP_P = 0.5
JITTER_P = 0.8
JITTER_BRIGHTNESS = 0.4
JITTER_CONTRAST = 0.4
JITTER_SATURATION = 0.2
JITTER_HUE = 0.1
GRAY_P = 0.2
U16_MAX = 65535


# ----------------------------------------------------------------------------------
# (1) chain restatement on top of torchvision (the reference's own dependency)
# ----------------------------------------------------------------------------------
class TwoViewChainTV:
    """Restates BYOLRGBDataTransforms (lightning_module.py:39-64) with torchvision v2.

    Same constructor arguments, same per-view op list and order, same ``__call__(x) ->
    [view1, view2]``.  ``x`` is a ``tv_tensors.Image`` float32 ``[C,H,W]`` in [0,1]
    (uint16 CPU tensors cannot run the chain: flip/ge are not implemented for UInt16,
    SURVEY F3), so callers pass ``u16.to(float32) * (1/65535)``.
    """

    def __init__(self, crop_size, mean, std, blur_prob=(1.0, 0.1), solarize_prob=(0.0, 0.2)):
        from torchvision.transforms import v2 as T

        if len(blur_prob) != 2 or len(solarize_prob) != 2:
            raise AssertionError("atm only 2 views are supported")
        self.crop_size = crop_size
        jitter = T.ColorJitter(JITTER_BRIGHTNESS, JITTER_CONTRAST, JITTER_SATURATION, JITTER_HUE)
        norm = T.Normalize(mean=mean, std=std)
        self.views = []
        for p_blur, p_sol in zip(blur_prob, solarize_prob):
            ops = [
                T.RandomResizedCrop(crop_size),                       # :49
                T.RandomHorizontalFlip(),                             # :50
                T.RandomApply([jitter], p=JITTER_P),                  # :51
                T.RandomGrayscale(p=GRAY_P),                          # :52
                T.RandomApply([T.GaussianBlur(kernel_size=23)], p=p_blur),  # :53
                T.RandomSolarize(128, p=p_sol),                       # :54
                T.ToImage(),                                          # :55
                T.ToDtype(torch.float32, scale=True),                 # :56
                norm,                                                 # :57
            ]
            self.views.append(T.Compose(ops))

    def __call__(self, x):
        return [view(x) for view in self.views]


class ResizeJitterChainTV:
    """Restates DecathlonDataModule.train_transforms / default_transforms (image branch,
    lightning_module.py:684-693, 703-711) with torchvision v2: Resize((s,s)) -> [ColorJitter(b, c)] -> ToDtype -> Normalize."""

    def __init__(self, size, mean, std, brightness=None, contrast=None):
        from torchvision.transforms import v2 as T
        ops = [T.ToImage(), T.Resize((size, size))]
        if brightness is not None or contrast is not None:
            ops.append(T.ColorJitter(brightness=brightness or 0, contrast=contrast or 0))
        ops += [T.ToDtype(torch.float32, scale=True), T.Normalize(mean=mean, std=std)]
        self.chain = T.Compose(ops)

    def __call__(self, x):
        return self.chain(x)


def u16_to_tv_image(x_u16: np.ndarray | torch.Tensor):
    """uint16 [H,W] or [C,H,W] -> tv_tensors.Image float32 [C,H,W] in [0,1]."""
    from torchvision import tv_tensors

    t = torch.as_tensor(np.asarray(x_u16).astype(np.int32)).to(torch.float32) * (1.0 / U16_MAX)
    if t.ndim == 2:
        t = t[None]
    return tv_tensors.Image(t)


# ----------------------------------------------------------------------------------
# (2a) RNG replay: the parameters of one view, in the reference's draw order (A.1)
# ----------------------------------------------------------------------------------
def _uniform(lo, hi) -> float:
    return torch.empty(1).uniform_(lo, hi).item()


def draw_view_params(H: int, W: int, blur_p: float = 0.0, solarize_p: float = 0.0) -> dict:
    """Consume torch's global generator exactly as one Compose of the chain does."""
    area = H * W
    log_ratio = torch.log(torch.tensor(RRC_RATIO))      # float32 bounds, _geometry.py:270
    box = None
    for _ in range(10):                                 # _geometry.py:277-292
        target_area = area * _uniform(RRC_SCALE[0], RRC_SCALE[1])
        aspect = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= W and 0 < h <= H:
            top = torch.randint(0, H - h + 1, size=(1,)).item()
            left = torch.randint(0, W - w + 1, size=(1,)).item()
            box = (top, left, h, w)
            break
    if box is None:                                     # central-crop fallback :293-306
        in_ratio = float(W) / float(H)
        if in_ratio < min(RRC_RATIO):
            w = W
            h = int(round(w / min(RRC_RATIO)))
        elif in_ratio > max(RRC_RATIO):
            h = H
            w = int(round(h * max(RRC_RATIO)))
        else:
            w, h = W, H
        box = ((H - h) // 2, (W - w) // 2, h, w)

    flip = not bool(torch.rand(1) >= FLIP_P)            # _transform.py:181
    jitter = not bool(torch.rand(1) >= JITTER_P)        # _container.py:104
    order = (0, 1, 2, 3)
    b = c = s = hue = 1.0
    hue = 0.0
    if jitter:                                          # _color.py:146-154
        order = tuple(int(v) for v in torch.randperm(4))
        b = _uniform(1 - JITTER_BRIGHTNESS, 1 + JITTER_BRIGHTNESS)
        c = _uniform(1 - JITTER_CONTRAST, 1 + JITTER_CONTRAST)
        s = _uniform(1 - JITTER_SATURATION, 1 + JITTER_SATURATION)
        hue = _uniform(-JITTER_HUE, JITTER_HUE)
    gray = not bool(torch.rand(1) >= GRAY_P)            # RandomGrayscale (identity at C=1)
    blur = not bool(torch.rand(1) >= blur_p)            # RandomApply([GaussianBlur])
    sigma = _uniform(0.1, 2.0) if blur else 0.0         # v2/_misc.py:209-211
    solarize = not bool(torch.rand(1) >= solarize_p)    # RandomSolarize
    return dict(top=box[0], left=box[1], h=box[2], w=box[3], flip=flip, jitter=jitter,
                order=order, brightness=b, contrast=c, saturation=s, hue=hue,
                gray=gray, blur=blur, sigma=sigma, solarize=solarize)


def draw_two_view_params(H, W, blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0)) -> list[dict]:
    """View 1 completely, then view 2 (lightning_module.py:63-64)."""
    return [draw_view_params(H, W, bp, sp) for bp, sp in zip(blur_prob, solarize_prob)]


# ----------------------------------------------------------------------------------
# (2b) antialiased bilinear resample (A.2)
# ----------------------------------------------------------------------------------
def aa_axis_weights(n: int, m: int):
    """Per-axis tap tables of ATen's antialiased bilinear filter, in_size n -> out_size m.

    Returns (xmin[m] int64, xsize[m] int64, weights[m, K] float32 zero-padded).
    """
    f32 = np.float32
    scale = f32(n) / f32(m)
    if scale >= 1.0:
        support = f32(scale)          # (interp_size / 2) * scale, interp_size == 2
        invscale = f32(1.0) / scale
    else:
        support = f32(1.0)
        invscale = f32(1.0)
    K = int(math.ceil(float(support))) * 2 + 1
    xmin = np.zeros(m, np.int64)
    xsize = np.zeros(m, np.int64)
    weights = np.zeros((m, K), np.float32)
    for i in range(m):
        center = f32(float(scale) * (i + 0.5))
        lo = max(int(float(center) - float(support) + 0.5), 0)
        hi = min(int(float(center) + float(support) + 0.5), n)
        size = min(max(hi - lo, 0), K)
        xmin[i], xsize[i] = lo, size
        total = f32(0.0)
        for j in range(size):
            arg = f32((f32(j + lo) - center + f32(0.5)) * invscale)
            wj = max(f32(0.0), f32(1.0) - abs(arg))
            weights[i, j] = wj
            total = f32(total + wj)
        if total != 0:
            weights[i, :size] = weights[i, :size] / total
    return xmin, xsize, weights


def aa_resize(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Separable AA bilinear, horizontal pass then vertical pass, float32 accumulate."""
    x = np.asarray(x, np.float32)
    h, w = x.shape
    xm, xs, xw = aa_axis_weights(w, out_w)
    tmp = np.zeros((h, out_w), np.float32)
    for i in range(out_w):
        acc = np.zeros(h, np.float32)
        for j in range(xs[i]):
            acc = acc + x[:, xm[i] + j] * xw[i, j]
        tmp[:, i] = acc
    ym, ys, yw = aa_axis_weights(h, out_h)
    out = np.zeros((out_h, out_w), np.float32)
    for i in range(out_h):
        acc = np.zeros(out_w, np.float32)
        for j in range(ys[i]):
            acc = acc + tmp[ym[i] + j, :] * yw[i, j]
        out[i, :] = acc
    return out


# ----------------------------------------------------------------------------------
# (2c) colour ops + normalise (A.3), and the whole view
# ----------------------------------------------------------------------------------
SOLARIZE_THRESHOLD = 128.0 / 255.0   # RandomSolarize(128) of lightning_module.py:54 on the [0,1] scale


def gaussian_kernel1d(sigma: float, kernel_size: int = 23) -> np.ndarray:
    """_get_gaussian_kernel1d (v2/functional/_misc.py:86-90) in float32: softmax(-(linspace(-lim, lim) / sigma)^2)."""
    f32 = np.float32
    lim = f32((kernel_size - 1) / (2.0 * math.sqrt(2.0)))
    step = f32((lim - (-lim)) / f32(kernel_size - 1))
    idx = np.arange(kernel_size)
    x = np.where(idx < kernel_size // 2, -lim + step * idx.astype(f32), lim - step * (kernel_size - 1 - idx).astype(f32)).astype(f32)
    v = -np.square((x / f32(sigma)).astype(f32)).astype(f32)
    e = np.exp(v - v.max()).astype(f32)
    return (e / e.sum(dtype=f32)).astype(f32)


def gaussian_blur(x: np.ndarray, sigma: float, kernel_size: int = 23) -> np.ndarray:
    """gaussian_blur_image (v2/functional/_misc.py:104-165) for one float plane: reflect padding by kernel_size // 2,
    then ONE 2-D convolution with the outer product of the 1-D kernels (float32)."""
    f32 = np.float32
    k1 = gaussian_kernel1d(sigma, kernel_size)
    k2 = (k1[:, None] * k1[None, :]).astype(f32)
    r = kernel_size // 2
    pad = np.pad(np.asarray(x, f32), r, mode="reflect")
    h, w = x.shape
    out = np.zeros((h, w), f32)
    for ky in range(kernel_size):
        for kx in range(kernel_size):
            out += k2[ky, kx] * pad[ky:ky + h, kx:kx + w]
    return out


def _gray(x: np.ndarray) -> np.ndarray:
    """_rgb_to_grayscale_image (functional/_color.py:31-48): r.mul(0.2989).add_(g, alpha=0.587).add_(b, alpha=0.114)."""
    f32 = np.float32
    return ((x[0] * f32(0.2989) + x[1] * f32(0.587)).astype(f32) + x[2] * f32(0.114)).astype(f32)


def _adjust_hue(x: np.ndarray, hue: float) -> np.ndarray:
    """adjust_hue_image for a float [3,h,w] image (functional/_color.py:300-396): _rgb_to_hsv, h = (h + hue) mod 1,
    _hsv_to_rgb -- the same expressions in float32."""
    f32 = np.float32
    r, g, b = x[0], x[1], x[2]
    maxc = x.max(axis=0)
    minc = x.min(axis=0)
    eqc = maxc == minc
    rng = (maxc - minc).astype(f32)
    sat = (rng / np.where(eqc, f32(1), maxc)).astype(f32)
    div = np.where(eqc, f32(1), rng).astype(f32)
    rc, gc, bc = (((maxc - c) / div).astype(f32) for c in (r, g, b))
    neq_r = maxc != r
    eq_g = maxc == g
    hg = ((rc + f32(2.0)).astype(f32) - bc).astype(f32) * (eq_g & neq_r)
    hr = (bc - gc).astype(f32) * (~neq_r)
    hb = ((gc + f32(4.0)).astype(f32) - rc).astype(f32) * (neq_r & ~eq_g)
    h = ((hr + hg).astype(f32) + hb).astype(f32)
    h = np.fmod((h * f32(1.0 / 6.0)).astype(f32) + f32(1.0), f32(1.0)).astype(f32)
    h = np.remainder((h + f32(hue)).astype(f32), f32(1.0)).astype(f32)
    h6 = (h * f32(6)).astype(f32)
    i = np.floor(h6)
    f = (h6 - i).astype(f32)
    i = np.remainder(i.astype(np.int32), 6)
    v = maxc
    sxf = (sat * f).astype(f32)
    oms = (f32(1.0) - sat).astype(f32)
    q = np.clip(((f32(1.0) - sxf).astype(f32) * v).astype(f32), 0, 1)
    t = np.clip(((sxf + oms).astype(f32) * v).astype(f32), 0, 1)
    p = np.clip((oms * v).astype(f32), 0, 1)
    vpqt = np.stack([v, p, q, t])
    select = np.array([[0, 2, 1, 1, 3, 0], [3, 0, 0, 2, 1, 1], [1, 1, 3, 0, 0, 2]])
    return np.take_along_axis(vpqt, select[:, i], axis=0).astype(f32)


def color_and_normalize_rgb(x: np.ndarray, params: dict, mean, std) -> np.ndarray:
    """The colour part of the chain for a float32 [3,s,s] image in [0,1] (lightning_module.py:51-57):
    ColorJitter ops in fn_idx order, RandomGrayscale, GaussianBlur, RandomSolarize, Normalize."""
    f32 = np.float32
    x = np.asarray(x, f32)
    if params["jitter"]:
        for k in params["order"]:
            if k == 0:
                x = np.clip(x * f32(params["brightness"]), f32(0), f32(1)).astype(f32)
            elif k == 1:      # mean over the grayscale image, functional/_color.py:199-204
                mu = _gray(x).mean(dtype=np.float64).astype(f32)
                c = float(params["contrast"])
                x = np.clip(x * f32(c) + mu * f32(1.0 - c), f32(0), f32(1)).astype(f32)
            elif k == 2:      # blend(x, gray(x), s), :151-166
                sf = float(params["saturation"])
                x = np.clip(x * f32(sf) + _gray(x)[None] * f32(1.0 - sf), f32(0), f32(1)).astype(f32)
            else:
                x = _adjust_hue(x, float(params["hue"]))
    if params.get("gray"):    # RandomGrayscale: gray replicated over the three channels (v2/_color.py:33-55)
        x = np.repeat(_gray(x)[None], 3, axis=0)
    if params.get("blur"):
        x = np.stack([gaussian_blur(x[c], params["sigma"]) for c in range(3)])
    if params.get("solarize"):
        x = np.where(x >= f32(SOLARIZE_THRESHOLD), f32(1) - x, x).astype(f32)
    m = np.asarray(mean, f32).reshape(3, 1, 1)
    sd = np.asarray(std, f32).reshape(3, 1, 1)
    return ((x - m) / sd).astype(f32)


def color_and_normalize(x: np.ndarray, params: dict, mean: float, std: float) -> np.ndarray:
    f32 = np.float32
    x = np.asarray(x, f32)
    if params["jitter"]:
        for k in params["order"]:
            if k == 0:      # adjust_brightness_image: x.mul(b).clamp_(0, 1)
                x = np.clip(x * f32(params["brightness"]), f32(0), f32(1)).astype(f32)
            elif k == 1:    # adjust_contrast_image -> _blend(x, mean(x), c)
                mu = x.mean(dtype=np.float64).astype(f32)
                c = float(params["contrast"])
                x = np.clip(x * f32(c) + mu * f32(1.0 - c), f32(0), f32(1)).astype(f32)
            # k == 2 (saturation) and k == 3 (hue) are identities when C == 1
            # (functional/_color.py:159-160, 380-381)
    # RandomGrayscale: identity when C == 1 (functional/_color.py:31-36)
    if params.get("blur"):          # RandomApply([GaussianBlur(23)]), lightning_module.py:53
        x = gaussian_blur(x, params["sigma"])
    if params.get("solarize"):      # RandomSolarize(128) on the [0,1] scale: x >= thr -> 1 - x (functional/_color.py:497-501)
        x = np.where(x >= f32(SOLARIZE_THRESHOLD), f32(1) - x, x).astype(f32)
    return ((x - f32(mean)) / f32(std)).astype(f32)


def apply_view(img_u16: np.ndarray, params: dict, crop_size: int, mean: float, std: float,
               window=(0.0, float(U16_MAX))) -> np.ndarray:
    """uint16 [H,W] + one view's parameters -> normalised float32 [s,s].

    ``window=(lo,hi)`` is the CT-windowing map clamp((x-lo)/(hi-lo),0,1); the identity
    default (0,65535) is the reference's ToDtype(scale=True) (functional/_misc.py:304).
    """
    f32 = np.float32
    lo, hi = window
    x = np.asarray(img_u16).astype(f32)
    if (lo, hi) == (0.0, float(U16_MAX)):
        x = x * f32(1.0 / U16_MAX)
    else:
        x = np.clip((x - f32(lo)) * f32(1.0 / (hi - lo)), f32(0), f32(1)).astype(f32)
    t, l, h, w = params["top"], params["left"], params["h"], params["w"]
    if x.ndim == 3:                                  # [3,H,W]: resample per channel, colour ops across channels
        out = np.stack([aa_resize(x[c, t:t + h, l:l + w], crop_size, crop_size) for c in range(x.shape[0])])
        if params["flip"]:
            out = out[:, :, ::-1]
        return color_and_normalize_rgb(out, params, mean, std)
    crop = x[t:t + h, l:l + w]                       # crop_image: slice before resize
    out = aa_resize(crop, crop_size, crop_size)
    if params["flip"]:
        out = out[:, ::-1]
    return color_and_normalize(out, params, mean, std)


def two_views(img_u16: np.ndarray, crop_size: int, mean: float, std: float, seed: int | None = None):
    """Restatement end-to-end for one image: returns ([v1, v2] float32 [s,s], [params1, params2])."""
    if seed is not None:
        torch.manual_seed(seed)
    H, W = img_u16.shape[-2:]
    ps = draw_two_view_params(H, W)
    return [apply_view(img_u16, p, crop_size, mean, std) for p in ps], ps
