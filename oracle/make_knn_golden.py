"""Generate tests/golden/knn.npz from the reference's own KNNOnlineEvaluator.predict (train/callback/knn.py:38-70).

    python -m oracle.make_knn_golden

The class is imported unmodified from /root/reference (stub modules for pytorch_lightning, oracle/ref_import.py) and run
on CPU in float32, as the callback would.  TEST INFRASTRUCTURE ONLY; needs /root/reference (build container).
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def load_reference_knn():
    if not ref_import.reference_available():
        raise RuntimeError("reference tree not found")
    ref_import._install_stubs()
    if ref_import.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_import.REFERENCE_ROOT)
    return importlib.import_module("medical_image_segmentation.train.callback.knn").KNNOnlineEvaluator


def make_case(seed: int, B: int, N: int, D: int, C: int, clustered: bool):
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, C, (N,), generator=g)
    if clustered:                                  # class centroids + noise: the evaluator's real regime
        cent = torch.randn(C, D, generator=g)
        bank = cent[labels] + 0.8 * torch.randn(N, D, generator=g)
        qlab = torch.randint(0, C, (B,), generator=g)
        query = cent[qlab] + 0.8 * torch.randn(B, D, generator=g)
    else:
        bank = torch.randn(N, D, generator=g)
        query = torch.randn(B, D, generator=g)
    bank = torch.nn.functional.normalize(bank, dim=1)          # knn.py:100
    query = torch.nn.functional.normalize(query, dim=1)        # knn.py:129
    return query, bank, labels


CASES = [("a", 1, 37, 500, 64, 10, 20, 0.1, True), ("b", 2, 64, 2000, 64, 100, 200, 0.07, True),
         ("c", 3, 16, 300, 32, 5, 300, 0.5, False), ("d", 4, 130, 1025, 96, 1000, 1, 0.07, True)]


def main():
    KNN = load_reference_knn()
    blob = {}
    for tag, seed, B, N, D, C, k, T, clustered in CASES:
        query, bank, labels = make_case(seed, B, N, D, C, clustered)
        ev = KNN(k=k, temperature=T, num_classes=C)
        pred = ev.predict(query, bank, labels)
        blob[f"{tag}_query"], blob[f"{tag}_bank"], blob[f"{tag}_labels"] = query.numpy(), bank.numpy(), labels.numpy()
        blob[f"{tag}_k"], blob[f"{tag}_T"], blob[f"{tag}_C"] = k, T, C
        blob[f"{tag}_pred"] = pred.numpy()[:, :5].astype(np.int32)      # the leading ranks (the callback uses [:, 0], :134);
                                                                        # far down, classes without a vote tie at score 0
    out = os.path.join(ROOT, "tests", "golden", "knn.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
