"""CPU restatement of the reference's weighted kNN prediction.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench CPU legs).

Follows KNNOnlineEvaluator.predict, /root/reference/medical_image_segmentation/train/callback/knn.py:38-70, step by step
in float64 numpy.  Pinned: tests/golden/knn.npz holds outputs of the reference's own method (oracle/make_knn_golden.py
imports the unmodified class) and tests/test_oracle_knn.py checks this restatement against them.
"""
from __future__ import annotations

import numpy as np


def knn_scores(query: np.ndarray, bank: np.ndarray, labels: np.ndarray, k: int, temperature: float, num_classes: int):
    """Returns (scores [B, C] float64, kth [B] = the k-th largest similarity, gap [B] = its distance to the (k+1)-th)."""
    q = np.asarray(query, np.float64)
    b = np.asarray(bank, np.float64)
    sim = q @ b.T                                                  # knn.py:52
    order = np.argsort(-sim, axis=1, kind="stable")                # knn.py:54 topk (ties: lower bank index first)
    idx = order[:, :k]
    w = np.exp(np.take_along_axis(sim, idx, axis=1) / temperature)  # knn.py:57
    lab = np.asarray(labels)[idx]                                  # knn.py:56
    scores = np.zeros((q.shape[0], num_classes), np.float64)       # knn.py:60-67: one-hot times weight, summed over k
    for r in range(q.shape[0]):
        np.add.at(scores[r], lab[r], w[r])
    kth = np.take_along_axis(sim, order[:, k - 1:k], axis=1)[:, 0]
    nxt = np.take_along_axis(sim, order[:, k:k + 1], axis=1)[:, 0] if k < b.shape[0] else np.full(q.shape[0], -np.inf)
    return scores, kth, kth - nxt


def knn_predict(query, bank, labels, k: int, temperature: float, num_classes: int) -> np.ndarray:
    """[B, C] classes by descending score (knn.py:70); equal scores by ascending class (torch leaves them unspecified)."""
    scores, _, _ = knn_scores(query, bank, labels, k, temperature, num_classes)
    return np.argsort(-scores, axis=1, kind="stable")
