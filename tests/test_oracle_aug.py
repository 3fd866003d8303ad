"""Pin the CPU oracle to the reference's own outputs (tests/golden, made by oracle/make_golden.py).

Reference path: BYOLRGBDataTransforms, train/data_loaders/lightning_module.py:39-64.
Gates: crop boxes / flips / jitter flags / op order bit-exact; factors exact
(same float32 draws); pixels |o - ref| <= 1e-3 * max(|ref|, 1)  (observed ~1e-6).
"""
import os

import numpy as np
import pytest
import torch

from oracle import aug_oracle as A
from oracle import ref_import
from tests import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PIX_TOL = 1e-3


def _check_params(ps, ints, order, fac):
    for p, i, o, f in zip(ps, ints, order, fac):
        assert (p["top"], p["left"], p["h"], p["w"]) == tuple(int(v) for v in i[:4])
        assert int(p["flip"]) == int(i[4]) and int(p["jitter"]) == int(i[5])
        if p["jitter"]:
            assert tuple(p["order"]) == tuple(int(v) for v in o)
            assert p["brightness"] == f[0] and p["contrast"] == f[1]
            assert p["saturation"] == f[2] and p["hue"] == f[3]


def _pix_ok(got, ref):
    return np.all(np.abs(got - ref) <= PIX_TOL * np.maximum(np.abs(ref), 1.0))


@pytest.mark.parametrize("crop", [32, 48])
def test_restatement_matches_reference_small(crop):
    g = np.load(os.path.join(GOLD, "aug_small.npz"))
    mean, std = float(g["mean"]), float(g["std"])
    for k, (img, seed) in enumerate(zip(g["images"], g["seeds"])):
        views, ps = A.two_views(img, crop, mean, std, seed=int(seed))
        _check_params(ps, g[f"ints_{crop}"][2 * k:2 * k + 2], g[f"order_{crop}"][2 * k:2 * k + 2],
                      g[f"fac_{crop}"][2 * k:2 * k + 2])
        for v in range(2):
            ref = g[f"out_{crop}"][k, v]
            assert _pix_ok(views[v], ref)
            assert np.abs(views[v] - ref).max() < 5e-6      # observed 1.2e-6


def test_restatement_matches_reference_real_slices():
    g = np.load(os.path.join(GOLD, "aug_real.npz"))
    mean, std = float(g["mean"]), float(g["std"])
    for k, (img, seed) in enumerate(zip(g["images"], g["seeds"])):
        views, ps = A.two_views(img, 64, mean, std, seed=int(seed))
        _check_params(ps, g["ints_64"][2 * k:2 * k + 2], g["order_64"][2 * k:2 * k + 2], g["fac_64"][2 * k:2 * k + 2])
        for v in range(2):
            assert np.abs(views[v] - g["out_64"][k, v]).max() < 5e-6


@pytest.mark.parametrize("crop", [224, 96, 256])
def test_restatement_matches_reference_512(crop):
    g = np.load(os.path.join(GOLD, "aug_512.npz"))
    mean, std = float(g["mean"]), float(g["std"])
    imgs = synth.batch_512(2)
    for k in range(2):
        views, ps = A.two_views(imgs[k], crop, mean, std, seed=int(g["seeds"][k]))
        _check_params(ps, g[f"ints_{crop}"][2 * k:2 * k + 2], g[f"order_{crop}"][2 * k:2 * k + 2],
                      g[f"fac_{crop}"][2 * k:2 * k + 2])
        for v in range(2):
            assert np.abs(views[v][::7, ::7] - g[f"sample_{crop}"][k, v]).max() < 5e-6
            assert abs(views[v].astype(np.float64).sum() - g[f"sum_{crop}"][k, v]) < 1e-2


def test_tv_chain_matches_reference_small():
    g = np.load(os.path.join(GOLD, "aug_small.npz"))
    chain = A.TwoViewChainTV(32, (float(g["mean"]),), (float(g["std"]),), (0.0, 0.0), (0.0, 0.0))
    for k, (img, seed) in enumerate(zip(g["images"], g["seeds"])):
        torch.manual_seed(int(seed))
        v1, v2 = chain(A.u16_to_tv_image(img))
        assert np.array_equal(v1[0].numpy(), g["out_32"][k, 0])
        assert np.array_equal(v2[0].numpy(), g["out_32"][k, 1])


@pytest.mark.parametrize("tag", ["512x512", "256x768", "448x448"])
def test_rng_replay_matches_reference_stream(tag):
    g = np.load(os.path.join(GOLD, "params_stream.npz"))
    H, W = (int(v) for v in tag.split("x"))
    torch.manual_seed(int(g["seed"]))
    ps = []
    for _ in range(400):
        ps.extend(A.draw_two_view_params(H, W))
    _check_params(ps, g[f"ints_{tag}"], g[f"order_{tag}"], g[f"fac_{tag}"])
    assert np.array_equal(torch.rand(4).numpy(), g[f"next_rand_{tag}"])   # same stream position


def test_aa_resize_matches_torch_interpolate():
    rng = np.random.default_rng(0)
    for (h, w, s) in ((300, 280, 224), (145, 190, 64), (64, 50, 96), (31, 200, 48)):
        x = rng.random((h, w), dtype=np.float32)
        ref = torch.nn.functional.interpolate(torch.from_numpy(x)[None, None], size=(s, s), mode="bilinear",
                                              align_corners=False, antialias=True)[0, 0].numpy()
        assert np.abs(A.aa_resize(x, s, s) - ref).max() < 1e-6


def test_window_identity_default():
    img = synth.uniform_slice(40, 48, 3)
    torch.manual_seed(5)
    p = A.draw_view_params(40, 48)
    a = A.apply_view(img, p, 16, 0.2, 0.3)
    b = A.apply_view(img, p, 16, 0.2, 0.3, window=(0.0, 65535.0))
    assert np.array_equal(a, b)


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree only exists in the build container")
def test_live_reference_agrees_with_golden():
    """Re-run the reference here and compare with the committed fixture (guards stale goldens)."""
    Ref = ref_import.load_reference_transforms()
    g = np.load(os.path.join(GOLD, "aug_small.npz"))
    chain = Ref(crop_size=48, mean=(float(g["mean"]),), std=(float(g["std"]),), blur_prob=(0.0, 0.0),
                solarize_prob=(0.0, 0.0))
    torch.manual_seed(int(g["seeds"][1]))
    v1, v2 = chain(A.u16_to_tv_image(g["images"][1]))
    assert np.array_equal(v1[0].numpy(), g["out_48"][1, 0])
    assert np.array_equal(v2[0].numpy(), g["out_48"][1, 1])


def test_restatement_matches_reference_golden_with_default_blur():
    """GaussianBlur(23) restated in numpy vs the reference class run with its default blur_prob=(1.0, 0.1)."""
    g = np.load(os.path.join(GOLD, "aug_blur.npz"))
    mean, std = float(g["mean"]), float(g["std"])
    worst = 0.0
    for crop in (32, 48):
        for k, seed in enumerate(g["seeds"]):
            torch.manual_seed(int(seed))
            ps = A.draw_two_view_params(96, 128, blur_prob=(1.0, 0.1), solarize_prob=(0.0, 0.0))
            for v in range(2):
                assert int(ps[v]["blur"]) == int(g[f"blur_{crop}"][2 * k + v])
                if ps[v]["blur"]:
                    assert ps[v]["sigma"] == g[f"sigma_{crop}"][2 * k + v]
                out = A.apply_view(g["images"][k], ps[v], crop, mean, std)
                worst = max(worst, float(np.abs(out - g[f"out_{crop}"][k, v]).max()))
    assert worst < 2e-5, worst


def test_native_replay_records_blur_and_solarize_draws():
    """csrc/rng_replay.cu records the GaussianBlur sigma and the RandomSolarize / RandomGrayscale outcomes it used to
    consume silently: identical to the torch replay and to the reference's recorded sigmas."""
    from medical_image_segmentation_b200 import params as P
    g = np.load(os.path.join(GOLD, "aug_blur.npz"))
    for k, seed in enumerate(g["seeds"]):
        torch.manual_seed(int(seed))
        p = P.draw_two_view_params(1, 96, 128, (1.0, 0.1), (0.0, 0.0))
        assert [int(f >> 3) & 1 for f in p["flags"]] == [int(b) for b in g["blur_32"][2 * k:2 * k + 2]]
        for v in range(2):
            if g["blur_32"][2 * k + v]:
                assert p["blur_sigma"][v] == np.float32(g["sigma_32"][2 * k + v])
    torch.manual_seed(5)
    a = P.draw_two_view_params_torch(200, 128, 160, (1.0, 0.1), (0.0, 0.2))
    torch.manual_seed(5)
    b = P.draw_two_view_params(200, 128, 160, (1.0, 0.1), (0.0, 0.2))
    assert a.tobytes() == b.tobytes()
    assert (a["flags"] & 16).any() and (a["flags"] & 4).any() and ((a["flags"] & 8) != 0)[::2].all()
    gen = torch.Generator().manual_seed(5)                  # a private generator draws the same stream
    before = torch.get_rng_state()
    c = P.draw_two_view_params(200, 128, 160, (1.0, 0.1), (0.0, 0.2), generator=gen)
    assert c.tobytes() == a.tobytes() and torch.equal(before, torch.get_rng_state())
