"""Weighted kNN prediction of the online evaluator on the GPU.

Mirrors ``KNNOnlineEvaluator`` (train/callback/knn.py:11-70): same constructor arguments (``k=200``,
``temperature=0.07``, ``num_classes=1000``, :28-36) and the same ``predict(query_feature, feature_bank, target_bank)``
-> ``[B, num_classes]`` class ranking (``pred_labels[:, 0]`` is the prediction, :131-134).  The Lightning callback around
it (feature-bank collection, ``concat_all_gather``, logging, :72-140) is orchestration and stays the reference's.

``sim = query @ bank.T`` runs on the tcgen05 GEMM of the loss path (hi/lo-split TF32 operands: fp32-grade similarities),
top-k / votes / ranking in one kernel per query
(csrc/knn.cu).  Equal scores (e.g. all the classes without a vote) are ranked by ascending class index, where torch's
``argsort`` leaves their order unspecified.  CUDA tensors only; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

launches = 0


class KNNOnlineEvaluator:
    def __init__(self, k: int = 200, temperature: float = 0.07, num_classes: int = 1000) -> None:
        self.num_classes = num_classes
        self.k = k
        self.temperature = temperature
        self._scratch = None

    def predict(self, query_feature: torch.Tensor, feature_bank: torch.Tensor, target_bank: torch.Tensor,
                return_scores: bool = False):
        """(B, D) queries, (N, D) bank, (N,) labels -> (B, num_classes) classes by descending weighted vote
        (and, with ``return_scores``, the (B, num_classes) scores)."""
        global launches
        if not (query_feature.is_cuda and feature_bank.is_cuda and target_bank.is_cuda):
            raise RuntimeError("KNNOnlineEvaluator.predict has no CPU path: features and labels must be CUDA tensors")
        if query_feature.dim() != 2 or feature_bank.dim() != 2 or query_feature.shape[1] != feature_bank.shape[1]:
            raise ValueError(f"expected (B, D) queries and an (N, D) bank, got {tuple(query_feature.shape)} and "
                             f"{tuple(feature_bank.shape)}")
        if target_bank.dim() != 1 or target_bank.shape[0] != feature_bank.shape[0]:
            raise ValueError(f"expected {feature_bank.shape[0]} bank labels, got {tuple(target_bank.shape)}")
        B, D = query_feature.shape
        N = feature_bank.shape[0]
        if not 1 <= self.k <= N:
            raise RuntimeError(f"selected index k out of range: k={self.k} for a bank of {N}")     # torch.topk's error
        dev = query_feature.device
        q = query_feature.detach().to(torch.float32).contiguous()
        bank = feature_bank.detach().to(torch.float32).contiguous()
        labels = target_bank.detach().to(torch.int64).contiguous()
        pred = torch.empty((B, self.num_classes), dtype=torch.int64, device=dev)
        scores = torch.empty((B, self.num_classes), dtype=torch.float32, device=dev) if return_scores else None
        if B == 0:
            return (pred, scores) if return_scores else pred
        need = int(_lib.lib.mis_knn_scratch_bytes(B, N, D))
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.empty((need,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib.mis_knn_predict(q.data_ptr(), bank.data_ptr(), labels.data_ptr(), B, N, D, int(self.k),
                                          1.0 / float(self.temperature), int(self.num_classes), pred.data_ptr(),
                                          scores.data_ptr() if return_scores else None, self._scratch.data_ptr(),
                                          self._scratch.numel(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "mis_knn_predict")
        launches += 4          # two operand splits, the GEMM, the vote kernel
        return (pred, scores) if return_scores else pred
