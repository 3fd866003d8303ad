"""B200-native hot path of EthanHaque/medical_image_segmentation's SSL pre-training step.

Public surface (mirrors the reference's transform / loss call sites):

    FusedTwoViewTransforms(crop_size, mean, std, blur_prob=(1.0,0.1), solarize_prob=(0.0,0.2))(x) -> [view1, view2]
    FusedFFCVTwoViewTransforms(device, crop_size, mean, std, solarize_prob).get_transforms() / .on_after_batch_transfer
    nt_xent_loss(z_a, z_b, temperature=0.1, group=None) -> scalar
    byol_cosine_loss(preds, targets) -> scalar
    compute_mean_and_std(loader) -> (mean, std)      (analyze_data/compute_dataset_metrics.py:12-29)
    momentum_update(online, momentum, m)             (BYOL.momentum_update, byol_pytorch.py:291-296; one kernel)
    KNNOnlineEvaluator(k, temperature, num_classes).predict(query, bank, labels)     (train/callback/knn.py:38-70)
    register_datamodule / get_datamodule            (registry hook of lightning_module.py:21-36)

Importing this package loads libmis_b200.so; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from . import peer  # noqa: F401
from .ema import momentum_update
from .knn import KNNOnlineEvaluator
from .loss import byol_cosine_loss, nt_xent_loss, nt_xent_rows
from .metrics import compute_mean_and_std
from .params import draw_two_view_params, draw_two_view_params_torch
from .registry import DATAMODULE_REGISTRY, get_datamodule, register_datamodule
from .transforms import (FusedFFCVTwoViewTransforms, FusedResizeJitterTransforms, FusedTwoViewTransforms,
                         algorithmic_bytes)

__all__ = [
    "FusedTwoViewTransforms", "FusedResizeJitterTransforms", "FusedFFCVTwoViewTransforms", "algorithmic_bytes", "nt_xent_loss", "nt_xent_rows", "byol_cosine_loss",
    "compute_mean_and_std", "momentum_update", "KNNOnlineEvaluator", "draw_two_view_params", "draw_two_view_params_torch", "register_datamodule", "get_datamodule",
    "DATAMODULE_REGISTRY",
]
