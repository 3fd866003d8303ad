"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (build container only).

    python -m oracle.make_golden            # from the repo root; needs /root/reference

Everything written here is produced by the unmodified reference classes imported from
/root/reference (oracle/ref_import.py): BYOLRGBDataTransforms
(train/data_loaders/lightning_module.py:39-64) and BYOL.cosine_similarity_loss
(train/model/byol_pytorch.py:181-198).  The oracle restatements and the CUDA path are
then tested against these files on any box (the GPU box has no /root/reference).

Files
-----
aug_small.npz   6 synthetic uint16 slices 96x128 (stored), crops 32 and 48: full outputs,
                recorded crop boxes / flip / jitter flags / op order / factors.
aug_real.npz    128x128 windows cut from the 5 real 16-bit slices shipped under
                data/visualizations/example_images/, crop 64: full outputs + params.
aug_512.npz     4 synthetic 512x512 slices (regenerated from seed by tests/synth.py),
                crops 224, 96 and 256: params, strided output samples out[::7, ::7], sums.
params_stream.npz  crop boxes / flags for 400 images (800 views) at 512x512, 256x768 and
                448x448 + the generator state after, pinning the RNG replay.
aug_blur.npz    the reference's DEFAULT blur_prob=(1.0, 0.1) (GaussianBlur(23) on view 1 always): 6 slices 96x128 at crops
                32 / 48 (full outputs, params, blur flags, sigmas) and 4 slices 512x512 at crop 224 (samples, sums).
aug_rgb.npz     8 synthetic 3-channel uint16 images 80x96 at crop 32 through the reference class (saturation, hue,
                RandomGrayscale live), without and with its default GaussianBlur.
byol_loss.npz   inputs and outputs of the reference BYOL loss.
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle.aug_oracle import u16_to_tv_image  # noqa: E402
from tests import synth  # noqa: E402

MEAN, STD = 0.227358, 0.237160     # lightning_module.py:212-213 rescaled from 0-255 to [0,1]
GOLD = os.path.join(ROOT, "tests", "golden")


class Recorder:
    """Hooks the reference's transform objects to record what they drew."""

    def __init__(self, chain):
        self.log = []
        for view_idx, compose in enumerate(chain.transforms):
            rrc, flip, jitter_apply = compose.transforms[0], compose.transforms[1], compose.transforms[2]
            jitter = jitter_apply.transforms[0]
            self._hook_params(rrc, view_idx, "box")
            self._hook_transform(flip, view_idx, "flip")
        # the ColorJitter object is shared by both views (lightning_module.py:44)
        self._hook_params(jitter, None, "jitter")

    def _hook_params(self, obj, view_idx, tag):
        orig = obj.make_params

        def wrapped(flat_inputs):
            p = orig(flat_inputs)
            self.log.append((tag, view_idx, p))
            return p

        obj.make_params = wrapped

    def _hook_transform(self, obj, view_idx, tag):
        orig = obj.transform

        def wrapped(inpt, params):
            self.log.append((tag, view_idx, None))
            return orig(inpt, params)

        obj.transform = wrapped

    def pop_views(self):
        """Split the log of one chain(x) call into two per-view records."""
        views, cur = [], None
        for tag, _, p in self.log:
            if tag == "box":
                cur = dict(top=p["top"], left=p["left"], h=p["height"], w=p["width"], flip=0, jitter=0,
                           order=(0, 1, 2, 3), brightness=1.0, contrast=1.0, saturation=1.0, hue=0.0)
                views.append(cur)
            elif tag == "flip":
                cur["flip"] = 1
            elif tag == "jitter":
                cur["jitter"] = 1
                cur["order"] = tuple(int(v) for v in p["fn_idx"])
                cur["brightness"] = p["brightness_factor"]
                cur["contrast"] = p["contrast_factor"]
                cur["saturation"] = p["saturation_factor"]
                cur["hue"] = p["hue_factor"]
        self.log.clear()
        assert len(views) == 2
        return views


def pack_params(views):
    """list of per-view dicts -> (int32 [n,6] top,left,h,w,flip,jitter ; int8 [n,4] ; float64 [n,4])."""
    ints = np.array([[v["top"], v["left"], v["h"], v["w"], v["flip"], v["jitter"]] for v in views], np.int32)
    order = np.array([v["order"] for v in views], np.int8)
    fac = np.array([[v["brightness"], v["contrast"], v["saturation"], v["hue"]] for v in views], np.float64)
    return ints, order, fac


def run_reference(Ref, images, crop, seeds):
    chain = Ref(crop_size=crop, mean=(MEAN,), std=(STD,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0))
    rec = Recorder(chain)
    outs, views = [], []
    for img, seed in zip(images, seeds):
        torch.manual_seed(int(seed))
        v1, v2 = chain(u16_to_tv_image(img))
        outs.append(np.stack([v1[0].numpy(), v2[0].numpy()]))
        views.extend(rec.pop_views())
    return np.stack(outs), pack_params(views)


def run_reference_blur(Ref, images, crop, seeds):
    """The reference class with its DEFAULT blur_prob=(1.0, 0.1) (lightning_module.py:40); solarize_prob=(0, 0) because
    RandomSolarize(128) raises on a float image (functional/_color.py:498-499).  Also records which views drew a blur
    and their sigma (GaussianBlur.make_params, v2/_misc.py:209-211)."""
    chain = Ref(crop_size=crop, mean=(MEAN,), std=(STD,), solarize_prob=(0.0, 0.0))
    rec = Recorder(chain)
    sig_log = []
    for view_idx, compose in enumerate(chain.transforms):
        blur = compose.transforms[4].transforms[0]
        orig = blur.make_params

        def wrapped(flat_inputs, _orig=orig, _v=view_idx):
            p = _orig(flat_inputs)
            sig_log.append((_v, p["sigma"][0]))
            return p

        blur.make_params = wrapped
    outs, views, blur_flag, sigma = [], [], [], []
    for img, seed in zip(images, seeds):
        torch.manual_seed(int(seed))
        sig_log.clear()
        v1, v2 = chain(u16_to_tv_image(img))
        outs.append(np.stack([v1[0].numpy(), v2[0].numpy()]))
        views.extend(rec.pop_views())
        got = dict(sig_log)
        for v in range(2):
            blur_flag.append(1 if v in got else 0)
            sigma.append(got.get(v, 0.0))
    return np.stack(outs), pack_params(views), np.array(blur_flag, np.int32), np.array(sigma, np.float64)


def make_blur_golden(Ref):
    """aug_blur.npz: the reference's default-argument chain (GaussianBlur on view 1 always, on view 2 with p = 0.1)."""
    imgs = np.stack([synth.ct_like_slice(96, 128, seed=70 + i) if i % 2 == 0
                     else synth.uniform_slice(96, 128, seed=70 + i) for i in range(6)])
    seeds = 3000 + np.arange(6)
    blob = dict(images=imgs, seeds=seeds, mean=MEAN, std=STD)
    for crop in (32, 48):
        out, (ints, order, fac), bf, sg = run_reference_blur(Ref, imgs, crop, seeds)
        blob[f"out_{crop}"] = out
        blob[f"ints_{crop}"], blob[f"order_{crop}"], blob[f"fac_{crop}"] = ints, order, fac
        blob[f"blur_{crop}"], blob[f"sigma_{crop}"] = bf, sg
    big = synth.batch_512(4)
    seeds512 = 1000 + np.arange(4)
    out, (ints, order, fac), bf, sg = run_reference_blur(Ref, big, 224, seeds512)
    blob["seeds_512"] = seeds512
    blob["sample_224"] = out[:, :, ::7, ::7].copy()
    blob["sum_224"] = out.astype(np.float64).sum(axis=(2, 3))
    blob["ints_224"], blob["blur_224"], blob["sigma_224"] = ints, bf, sg
    np.savez_compressed(os.path.join(GOLD, "aug_blur.npz"), **blob)


def make_rgb_golden(Ref):
    """aug_rgb.npz: the reference class on 3-channel float input (what its IMAGENET / CIFAR modules feed it after
    decoding, lightning_module.py:408,482-488, here from uint16 RGB so that no 8-bit rounding enters): saturation, hue,
    RandomGrayscale and the default GaussianBlur(23); solarize_prob=(0, 0) (RandomSolarize(128) raises on float images)."""
    imgs = np.stack([np.stack([synth.ct_like_slice(80, 96, seed=90 + 3 * i + c) if (i + c) % 2 == 0
                               else synth.uniform_slice(80, 96, seed=90 + 3 * i + c) for c in range(3)]) for i in range(8)])
    seeds = 5000 + np.arange(8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    blob = dict(images=imgs, seeds=seeds, mean=np.array(mean), std=np.array(std))
    for tag, kw in (("noblur", dict(blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0))), ("blur", dict(solarize_prob=(0.0, 0.0)))):
        chain = Ref(crop_size=32, mean=mean, std=std, **kw)
        outs = []
        for img, seed in zip(imgs, seeds):
            torch.manual_seed(int(seed))
            v1, v2 = chain(u16_to_tv_image(img))
            outs.append(np.stack([v1.numpy(), v2.numpy()]))
        blob[f"out_{tag}"] = np.stack(outs)
    np.savez_compressed(os.path.join(GOLD, "aug_rgb.npz"), **blob)


def main():
    import cv2
    if "--blur-only" in sys.argv:
        make_blur_golden(ref_import.load_reference_transforms())
        return
    if "--rgb-only" in sys.argv:
        make_rgb_golden(ref_import.load_reference_transforms())
        return

    Ref = ref_import.load_reference_transforms()
    os.makedirs(GOLD, exist_ok=True)

    # ---- aug_small ---------------------------------------------------------------
    imgs = np.stack([synth.ct_like_slice(96, 128, seed=50 + i) if i % 2 == 0
                     else synth.uniform_slice(96, 128, seed=50 + i) for i in range(6)])
    seeds = 1000 + np.arange(6)
    blob = dict(images=imgs, seeds=seeds, mean=MEAN, std=STD)
    for crop in (32, 48):
        out, (ints, order, fac) = run_reference(Ref, imgs, crop, seeds)
        blob[f"out_{crop}"] = out
        blob[f"ints_{crop}"], blob[f"order_{crop}"], blob[f"fac_{crop}"] = ints, order, fac
    np.savez_compressed(os.path.join(GOLD, "aug_small.npz"), **blob)

    # ---- aug_real: windows of the shipped 16-bit slices -----------------------------
    files = sorted(glob.glob(os.path.join(ref_import.REFERENCE_ROOT, "data", "visualizations",
                                          "example_images", "*.png")))
    wins = []
    for f in files:
        im = cv2.imread(f, cv2.IMREAD_UNCHANGED)
        assert im.dtype == np.uint16
        h0, w0 = (im.shape[0] - 128) // 2, (im.shape[1] - 128) // 2
        wins.append(im[h0:h0 + 128, w0:w0 + 128].copy())
    wins = np.stack(wins)
    seeds = 2000 + np.arange(len(wins))
    out, (ints, order, fac) = run_reference(Ref, wins, 64, seeds)
    np.savez_compressed(os.path.join(GOLD, "aug_real.npz"), images=wins, seeds=seeds, mean=MEAN, std=STD,
                        out_64=out, ints_64=ints, order_64=order, fac_64=fac,
                        names=np.array([os.path.basename(f)[:16] for f in files]))

    # ---- aug_512: full-size synthetic slices, sampled outputs -------------------------
    big = synth.batch_512(4)
    seeds = 1000 + np.arange(4)
    blob = dict(seeds=seeds, mean=MEAN, std=STD, stride=7)
    for crop in (224, 96, 256):
        out, (ints, order, fac) = run_reference(Ref, big, crop, seeds)
        blob[f"sample_{crop}"] = out[:, :, ::7, ::7].copy()
        blob[f"sum_{crop}"] = out.astype(np.float64).sum(axis=(2, 3))
        blob[f"sumsq_{crop}"] = (out.astype(np.float64) ** 2).sum(axis=(2, 3))
        blob[f"ints_{crop}"], blob[f"order_{crop}"], blob[f"fac_{crop}"] = ints, order, fac
    np.savez_compressed(os.path.join(GOLD, "aug_512.npz"), **blob)

    make_blur_golden(Ref)
    make_rgb_golden(Ref)

    # ---- params_stream: RNG replay pin ----------------------------------------------
    blob = {}
    for (H, W) in ((512, 512), (256, 768), (448, 448)):
        chain = Ref(crop_size=8, mean=(MEAN,), std=(STD,), blur_prob=(0.0, 0.0), solarize_prob=(0.0, 0.0))
        rec = Recorder(chain)
        x = u16_to_tv_image(np.zeros((H, W), np.uint16))
        torch.manual_seed(4242)
        views = []
        for _ in range(400):
            chain(x)
            views.extend(rec.pop_views())
        ints, order, fac = pack_params(views)
        tag = f"{H}x{W}"
        blob[f"ints_{tag}"], blob[f"order_{tag}"], blob[f"fac_{tag}"] = ints, order, fac
        blob[f"next_rand_{tag}"] = torch.rand(4).numpy()      # stream position after 400 images
    np.savez_compressed(os.path.join(GOLD, "params_stream.npz"), seed=4242, **blob)

    # ---- byol_loss ------------------------------------------------------------------
    loss_fn = ref_import.load_reference_byol_loss()
    g = torch.Generator().manual_seed(7)
    cases = {}
    for i, (n, d) in enumerate(((8, 16), (64, 256), (33, 128))):
        p = torch.randn(n, d, generator=g)
        t = torch.randn(n, d, generator=g)
        if i == 2:
            t[3] = 0.0          # exercises the eps clamp of F.normalize
        cases[f"preds_{i}"], cases[f"targets_{i}"] = p.numpy(), t.numpy()
        cases[f"loss_{i}"] = np.float32(loss_fn(p, t).item())
    np.savez_compressed(os.path.join(GOLD, "byol_loss.npz"), **cases)

    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
