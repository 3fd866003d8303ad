"""Import the reference's own hot-path classes (build container only; TEST INFRASTRUCTURE).

/root/reference exists only in the build container, never on the GPU box, so this
module is used exclusively by oracle/make_golden.py (fixture generation) and by the
CPU tests that are skipped when the reference is absent.

medical_image_segmentation/train/data_loaders/lightning_module.py imports
pytorch_lightning (:3,6), ffcv (:7-13) and -- through analyze_data/pytorch_datasets.py:9
-- matplotlib; none is installed.  None of them is touched by BYOLRGBDataTransforms or
BYOL.cosine_similarity_loss, so empty stub modules are enough.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MIS_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.utilities.types",
    "pytorch_lightning.callbacks", "pytorch_lightning.loggers",
    "ffcv", "ffcv.fields", "ffcv.fields.decoders", "ffcv.transforms", "ffcv.loader",
    "ffcv.pipeline", "ffcv.pipeline.operation",
    "matplotlib", "matplotlib.pyplot",
    "pydicom", "nibabel", "cv2_stub_unused",
]


class _Anything:
    """Attribute sink: any name resolves to a dummy class usable as a base class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "medical_image_segmentation"))


def _install_stubs():
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        sink = _Anything()
        mod.__getattr__ = sink.__getattr__          # PEP 562 module-level __getattr__
        mod.__path__ = []                            # behave like a package
        sys.modules[name] = mod


def load_reference_transforms():
    """Return the reference's BYOLRGBDataTransforms class (lightning_module.py:39-64)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mod = importlib.import_module("medical_image_segmentation.train.data_loaders.lightning_module")
    return mod.BYOLRGBDataTransforms


def load_reference_byol_loss():
    """Return BYOL.cosine_similarity_loss as an unbound function (byol_pytorch.py:181-198)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mod = importlib.import_module("medical_image_segmentation.train.model.byol_pytorch")
    fn = mod.BYOL.cosine_similarity_loss
    return lambda preds, targets: fn(None, preds, targets)
