"""Data-module registry hook with the reference's names (train/data_loaders/lightning_module.py:21-36),
plus the one module this package owns: ``SYNTH_U16`` -- raw uint16 slices resident on the GPU, augmented
per batch by the fused kernel (what bench.py and the tests drive)."""
from __future__ import annotations

import torch

from .transforms import FusedTwoViewTransforms

DATAMODULE_REGISTRY = {}


def register_datamodule(name):
    def decorator(cls):
        DATAMODULE_REGISTRY[name] = cls
        return cls

    return decorator


def get_datamodule(name):
    if name in DATAMODULE_REGISTRY:
        return DATAMODULE_REGISTRY[name]
    raise ValueError(f"No datamodule registered with name {name}")


@register_datamodule("SYNTH_U16")
class SyntheticU16DataModule:
    """Yields ``(view_1, label, view_2)`` batches like the reference's loaders (byol_pytorch.py:201-204).

    mean/std default to the RADIOLOGY constants (lightning_module.py:212-213: 57.9764 / 60.4759 on the
    0-255 scale) rescaled to [0,1].
    """
    mean = (57.9764 / 255.0,)
    std = (60.4759 / 255.0,)

    def __init__(self, batch_size=256, image_size=512, crop_size=112, num_batches=8, device="cuda", seed=1234,
                 out_dtype=torch.bfloat16):
        self.batch_size, self.image_size, self.num_batches = batch_size, image_size, num_batches
        self.device, self.seed = device, seed
        self.transforms = FusedTwoViewTransforms(crop_size, self.mean, self.std, out_dtype=out_dtype)

    def train_dataloader(self):
        g = torch.Generator(device=self.device).manual_seed(self.seed)
        for _ in range(self.num_batches):
            x = torch.randint(0, 65536, (self.batch_size, 1, self.image_size, self.image_size), dtype=torch.int32,
                              device=self.device, generator=g).to(torch.uint16)
            labels = torch.zeros(self.batch_size, dtype=torch.long, device=self.device)
            v1, v2 = self.transforms(x)
            yield v1, labels, v2
