"""Host RNG replay (csrc/rng_replay.cu) is bit-exact with the reference's parameter stream.

Golden: tests/golden/params_stream.npz, recorded from the reference's own transform objects
(torchvision RandomResizedCrop / RandomHorizontalFlip / ColorJitter hooked by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from medical_image_segmentation_b200 import params as P
from medical_image_segmentation_b200._lib import VIEW_PARAMS_DTYPE

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag", ["512x512", "256x768", "448x448"])
def test_native_replay_matches_reference_stream(tag):
    g = np.load(os.path.join(GOLD, "params_stream.npz"))
    H, W = (int(v) for v in tag.split("x"))
    torch.manual_seed(int(g["seed"]))
    p = P.draw_two_view_params(400, H, W)
    ints = g[f"ints_{tag}"]
    assert np.array_equal(np.stack([p["top"], p["left"], p["h"], p["w"]], 1), ints[:, :4])
    assert np.array_equal(p["flags"] & 1, ints[:, 4])
    assert np.array_equal((p["flags"] >> 1) & 1, ints[:, 5])
    jit = ints[:, 5] == 1
    assert np.array_equal(p["order"][jit], g[f"order_{tag}"][jit].astype(np.uint8))
    fac = g[f"fac_{tag}"]
    for k, name in enumerate(("brightness", "contrast", "saturation", "hue")):
        assert np.array_equal(p[name][jit].astype(np.float64), fac[jit, k])      # same float32 draws
    assert np.array_equal(p["img"], np.repeat(np.arange(400), 2))
    assert np.array_equal(torch.rand(4).numpy(), g[f"next_rand_{tag}"])          # generator left where torch leaves it


@pytest.mark.parametrize("shape", [(512, 512), (64, 700), (1000, 90), (33, 33)])
def test_native_replay_equals_torch_replay(shape):
    H, W = shape
    torch.manual_seed(99)
    a = P.draw_two_view_params_torch(300, H, W)
    nxt = torch.rand(2)
    torch.manual_seed(99)
    b = P.draw_two_view_params(300, H, W)
    for name in VIEW_PARAMS_DTYPE.names:
        assert np.array_equal(a[name], b[name]), name
    assert torch.equal(nxt, torch.rand(2))


def test_blur_probability_consumes_sigma_draw():
    torch.manual_seed(5)
    a = P.draw_two_view_params_torch(50, 128, 128, blur_prob=(1.0, 0.1))
    torch.manual_seed(5)
    b = P.draw_two_view_params(50, 128, 128, blur_prob=(1.0, 0.1))
    for name in VIEW_PARAMS_DTYPE.names:
        assert np.array_equal(a[name], b[name]), name


def test_generator_state_crosses_mt_block_boundary():
    """> 624 words per call forces mt19937 state regeneration inside the native code."""
    torch.manual_seed(1)
    a = P.draw_two_view_params(2000, 512, 512)
    torch.manual_seed(1)
    b = np.concatenate([P.draw_two_view_params(n, 512, 512) for n in (1, 7, 992, 1000)])
    b["img"] = np.repeat(np.arange(2000), 2)
    for name in VIEW_PARAMS_DTYPE.names:
        assert np.array_equal(a[name], b[name]), name


def test_view_major_reorder():
    from medical_image_segmentation_b200 import FusedTwoViewTransforms
    p = np.zeros(6, VIEW_PARAMS_DTYPE)
    p["top"] = np.arange(6)
    q = FusedTwoViewTransforms.to_view_major(p)
    assert list(q["top"]) == [0, 2, 4, 1, 3, 5]


def _handback_draw(n_images, H, W):
    """The hand-back protocol of mis_draw_two_view_params (no callback): native until an image's box depends on the
    last bit of torch.exp, that image by the pure-torch replay, then native again.  Returns (records, handed-back images)."""
    import ctypes as C
    from medical_image_segmentation_b200 import _lib
    out = np.zeros(2 * n_images, VIEW_PARAMS_DTYPE)
    zero = (C.c_float * 2)(0.0, 0.0)
    n_done = C.c_int(0)
    state = torch.get_rng_state()
    blob = state.numpy()
    done, handed = 0, []
    while done < n_images:
        rc = _lib.lib.mis_draw_two_view_params(blob.ctypes.data, blob.nbytes, n_images - done, done, H, W,
                                               C.cast(zero, C.c_void_p), C.cast(zero, C.c_void_p),
                                               out[2 * done:].ctypes.data, C.byref(n_done))
        assert rc == 0
        done += n_done.value
        if done < n_images:
            torch.set_rng_state(state)
            out[2 * done:2 * done + 2] = P.draw_two_view_params_torch(1, H, W, img0=done)
            handed.append(done)
            state = torch.get_rng_state()
            blob = state.numpy()
            done += 1
    torch.set_rng_state(state)
    return out, handed


def test_exp_sensitive_images_stay_bit_exact(monkeypatch):
    """~2.6e-4 of the images have a crop box within rounding distance of a .5 boundary, where the last bit of the
    reference's float32 torch.exp decides.  Two protocols: the product path asks torch for that one value through a
    callback and never leaves native code; the plain ABI call hands the image back and torch draws it.  Both must give
    the records of the pure-torch replay and leave the stream aligned."""
    calls = []
    real = P._torch_exp_f32

    torch.manual_seed(123)
    legacy, handed = _handback_draw(30000, 512, 512)
    after_legacy = torch.rand(3)
    assert len(handed) >= 1
    torch.manual_seed(123)
    p = P.draw_two_view_params(30000, 512, 512)                    # callback protocol
    assert torch.equal(after_legacy, torch.rand(3))
    assert p.tobytes() == legacy.tobytes()
    i = handed[0]
    torch.manual_seed(123)
    P.draw_two_view_params(i, 512, 512)          # position the stream at image i
    ref = P.draw_two_view_params_torch(3, 512, 512, img0=i)        # pure torch replay of image i and its two successors
    for name in VIEW_PARAMS_DTYPE.names:
        assert np.array_equal(ref[name], p[name][2 * i:2 * i + 6]), name
    # a rewind across a generator refill (624 words): hand-backs at every position of the state array
    for seed in range(40):
        torch.manual_seed(1000 + seed)
        a, _ = _handback_draw(4000, 512, 512)
        torch.manual_seed(1000 + seed)
        assert P.draw_two_view_params(4000, 512, 512).tobytes() == a.tobytes(), seed


def test_resize_jitter_params_match_torchvision_colorjitter():
    """Decathlon flavour (lightning_module.py:689): ColorJitter(brightness=0.2, contrast=0.2).make_params stream."""
    from torchvision.transforms import v2 as T
    from medical_image_segmentation_b200 import FusedResizeJitterTransforms
    cj = T.ColorJitter(brightness=0.2, contrast=0.2)
    torch.manual_seed(11)
    ref = [cj.make_params([]) for _ in range(300)]
    nxt = torch.rand(2)
    t = FusedResizeJitterTransforms(224, (0.1,), (0.2,), brightness=0.2, contrast=0.2)
    torch.manual_seed(11)
    p = t.draw_params(300, 320, 320)
    assert torch.equal(nxt, torch.rand(2))
    for i, r in enumerate(ref):
        assert tuple(int(v) for v in r["fn_idx"]) == tuple(int(v) for v in p["order"][i])
        assert np.float32(r["brightness_factor"]) == p["brightness"][i]
        assert np.float32(r["contrast_factor"]) == p["contrast"][i]
        assert r["saturation_factor"] is None and r["hue_factor"] is None
    assert np.all(p["h"] == 320) and np.all(p["top"] == 0) and np.all(p["flags"] == 2)
    # default_transforms: no jitter, no RNG consumption
    t0 = FusedResizeJitterTransforms(224, (0.1,), (0.2,))
    torch.manual_seed(3)
    a = torch.rand(1)
    torch.manual_seed(3)
    p0 = t0.draw_params(5, 64, 64)
    assert torch.equal(a, torch.rand(1)) and np.all(p0["flags"] == 0)


def test_prefetched_parameter_stream_is_the_same_stream():
    """prefetch_params=True draws batch k+1 on a helper thread while batch k is in flight: same records, same order."""
    from medical_image_segmentation_b200 import FusedTwoViewTransforms
    a = FusedTwoViewTransforms(64, (0.2,), (0.2,))
    b = FusedTwoViewTransforms(64, (0.2,), (0.2,), prefetch_params=True)
    torch.manual_seed(99)
    ref = [a.draw_params(37, 256, 320) for _ in range(5)] + [a.draw_params(8, 128, 128)]
    torch.manual_seed(99)
    got = [b.next_params(37, 256, 320) for _ in range(5)]
    b.drain_prefetch()
    # the helper thread had already drawn a 6th (37, 256, 320) batch: a shape change redraws from the generator as it
    # stands, so only the first five batches are comparable one to one
    for r, g in zip(ref[:5], got):
        assert r.tobytes() == g.tobytes()
    torch.manual_seed(5)
    x = a.draw_params(4, 64, 64)
    torch.manual_seed(5)
    assert b.next_params(4, 64, 64).tobytes() == x.tobytes()
    b.drain_prefetch()
    # view_major=True: the helper thread also reorders; switching the order mid-stream keeps the same draws
    torch.manual_seed(17)
    ref = [a.draw_params(6, 96, 80) for _ in range(4)]
    torch.manual_seed(17)
    got = [b.next_params(6, 96, 80, view_major=True), b.next_params(6, 96, 80, view_major=True),
           b.next_params(6, 96, 80), b.next_params(6, 96, 80, view_major=True)]
    b.drain_prefetch()
    for k, (r, g) in enumerate(zip(ref, got)):
        want = r if k == 2 else a.to_view_major(r)
        assert want.tobytes() == g.tobytes(), k


def test_ffcv_flavour_parameter_records():
    """BYOLRGBFFCVDataTransforms' op set (lightning_module.py:77-95): crop, flip, grayscale, solarize -- no colour jitter,
    no blur.  Same crop boxes / flips as the torchvision-flavour stream for the same seed."""
    from medical_image_segmentation_b200 import FusedFFCVTwoViewTransforms, FusedTwoViewTransforms
    f = FusedFFCVTwoViewTransforms("cuda:0", (112, 112), (0.2,), (0.2,))
    assert f.crop_size == 112 and f.blur_prob == (0.0, 0.0) and f.solarize_prob == (0.0, 0.2)
    assert len(f.get_transforms()) == 2 and all(len(p) == 1 and callable(p[0]) for p in f.get_transforms())
    torch.manual_seed(9)
    a = f.draw_params(300, 256, 256)
    torch.manual_seed(9)
    b = FusedTwoViewTransforms(112, (0.2,), (0.2,), blur_prob=(0.0, 0.0)).draw_params(300, 256, 256)
    assert not (a["flags"] & (2 | 8)).any() and (a["flags"] & 16).any() and (a["flags"] & 4).any()
    for k in ("img", "top", "left", "h", "w"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["flags"] & 1, b["flags"] & 1) and np.array_equal(a["flags"] & 16, b["flags"] & 16)


def test_view_cost_order_is_a_stable_descending_permutation():
    """mis_view_cost_order: most expensive (largest crop area) views first, ties in table order."""
    import ctypes as C
    from medical_image_segmentation_b200 import FusedTwoViewTransforms, _lib
    t = FusedTwoViewTransforms(64, (0.2,), (0.2,))
    torch.manual_seed(4)
    p = np.ascontiguousarray(t.to_view_major(t.draw_params(300, 512, 512)))
    order = np.empty(p.shape[0], np.int32)
    assert _lib.lib.mis_view_cost_order(p.ctypes.data, p.shape[0], order.ctypes.data) == 0
    assert np.array_equal(np.sort(order), np.arange(p.shape[0]))
    area = p["h"].astype(np.int64) * p["w"]
    bucket = 1023 - area * 1023 // area.max()
    b = bucket[order]
    assert (np.diff(b) >= 0).all()                                   # non-increasing area at the sort's resolution
    same = np.diff(b) == 0
    assert (np.diff(order)[same] > 0).all()                          # stable inside a bucket
    assert _lib.lib.mis_view_cost_order(None, 0, None) == 0


def test_native_table_check_reports_the_first_bad_record():
    """mis_view_params_check (what apply() runs before the kernel trusts a table): boxes outside the slice, slices
    outside the batch and a blurred view without a usable sigma are reported by index; flags are OR-ed up to there."""
    import ctypes as C
    from medical_image_segmentation_b200 import FusedTwoViewTransforms, _lib
    t = FusedTwoViewTransforms(64, (0.2,), (0.2,), blur_prob=(1.0, 0.1), solarize_prob=(0.0, 0.2))
    torch.manual_seed(8)
    p = np.ascontiguousarray(t.to_view_major(t.draw_params(50, 256, 320)))

    def check(tab, B=50, H=256, W=320):
        flags, bad = C.c_uint32(0), C.c_int(-2)
        assert _lib.lib.mis_view_params_check(tab.ctypes.data, tab.shape[0], B, H, W, C.byref(flags), C.byref(bad)) == 0
        return flags.value, bad.value

    flags, bad = check(p)
    assert bad == -1 and flags == int(np.bitwise_or.reduce(p["flags"])) and (flags & 8)
    for field, value, k in (("top", 256, 7), ("left", -1, 0), ("h", 0, 99), ("w", 321, 42), ("img", 50, 13), ("img", -1, 60)):
        q = p.copy()
        q[field][k] = value
        assert check(q)[1] == k, (field, value)
    q = p.copy()
    k = int(np.flatnonzero(q["flags"] & 8)[3])
    q["blur_sigma"][k] = 0.0
    assert check(q)[1] == k
    q["blur_sigma"][k] = np.inf
    assert check(q)[1] == k
    assert check(p, B=49)[1] == int(np.flatnonzero(p["img"] >= 49)[0])
    assert check(p[:0]) == (0, -1)
    # the Python face of it
    with pytest.raises(ValueError):
        q = p.copy()
        q["top"][3] = 250
        FusedTwoViewTransforms._validate_table(q, 50, 256, 320)
