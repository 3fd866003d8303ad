"""Weighted kNN prediction on the GPU (csrc/knn.cu via KNNOnlineEvaluator.predict) against the reference's own outputs
(tests/golden/knn.npz, written by oracle/make_knn_golden.py from the unmodified reference class) and the fp64 oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import knn_oracle as K

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _run(query, bank, labels, k, T, C):
    from medical_image_segmentation_b200 import KNNOnlineEvaluator
    ev = KNNOnlineEvaluator(k=k, temperature=T, num_classes=C)
    pred, scores = ev.predict(torch.from_numpy(query).cuda(), torch.from_numpy(bank).cuda(), torch.from_numpy(labels).cuda(),
                              return_scores=True)
    assert pred.shape == (query.shape[0], C) and pred.dtype == torch.int64
    return pred.cpu().numpy(), scores.cpu().numpy().astype(np.float64)


def _check(query, bank, labels, k, T, C, ref_top=None):
    pred, scores = _run(query, bank, labels, k, T, C)
    want, kth, gap = K.knn_scores(query, bank, labels, k, T, C)
    # a neighbour swap at the k-th place is legitimate when the k-th and (k+1)-th similarities differ by less than fp32
    # resolution; such rows are compared through the ranking only
    clean = gap > 1e-6
    err = np.abs(scores - want).max(axis=1) / np.maximum(want.max(axis=1), 1e-30)
    assert err[clean].max() <= 1e-3, err[clean].max()          # north_star's floating-point gate (observed ~1e-6)
    # every row of pred is a permutation of the classes, ordered by descending score, equal scores by ascending class
    assert np.array_equal(np.sort(pred, axis=1), np.tile(np.arange(C), (pred.shape[0], 1)))
    ranked = np.take_along_axis(scores, pred, axis=1)
    assert (np.diff(ranked, axis=1) <= 0).all()
    same = np.diff(ranked, axis=1) == 0
    assert (np.diff(pred, axis=1)[same] > 0).all()
    # the prediction itself: identical to the oracle's wherever the two leading scores are not a tie
    wpred = np.argsort(-want, axis=1, kind="stable")
    top2 = np.sort(want, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 1e-4 * top2[:, 1]
    assert np.array_equal(pred[decided & clean, 0], wpred[decided & clean, 0])
    if ref_top is not None:                                    # the reference's own float32 run
        agree = pred[:, 0] == ref_top[:, 0]
        assert agree[decided & clean].all() and agree.mean() >= 0.98


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_matches_reference_outputs(tag):
    g = np.load(os.path.join(GOLD, "knn.npz"))
    _check(g[f"{tag}_query"], g[f"{tag}_bank"], g[f"{tag}_labels"], int(g[f"{tag}_k"]), float(g[f"{tag}_T"]),
           int(g[f"{tag}_C"]), g[f"{tag}_pred"])


def test_large_bank_and_ties():
    """A bank larger than one GEMM tile row, with duplicated bank rows (exact similarity ties, some straddling the k-th
    place) and k at the kernel's maximum."""
    gen = torch.Generator().manual_seed(5)
    bank = torch.nn.functional.normalize(torch.randn(20000, 128, generator=gen), dim=1)
    bank[10000:] = bank[:10000]                      # every similarity occurs twice
    labels = torch.randint(0, 37, (20000,), generator=gen)
    query = torch.nn.functional.normalize(torch.randn(200, 128, generator=gen), dim=1)
    for k in (1, 7, 1024):
        pred, scores = _run(query.numpy(), bank.numpy(), labels.numpy(), k, 0.1, 37)
        want, _, _ = K.knn_scores(query.numpy(), bank.numpy(), labels.numpy(), k, 0.1, 37)   # stable: lower index wins ties
        err = np.abs(scores - want).max(axis=1) / want.max(axis=1)
        # duplicates have bit-identical similarities on the GPU too (same products, same order), so the lower-index rule
        # selects the same neighbours as the oracle
        assert np.quantile(err, 0.99) <= 1e-3 and (err <= 1e-3).mean() >= 0.98, (k, err.max())


def test_masses_of_equal_similarities_take_the_radix_path():
    """6000 identical bank rows (more than the fast path's candidate list holds): every similarity ties, the k lowest
    bank indices are the neighbours, exactly as the oracle's stable ordering selects them."""
    gen = torch.Generator().manual_seed(9)
    base = torch.nn.functional.normalize(torch.randn(1, 64, generator=gen), dim=1)
    bank = base.repeat(6000, 1)
    bank[::1000] = torch.nn.functional.normalize(torch.randn(6, 64, generator=gen), dim=1)   # a few distinct rows
    labels = torch.arange(6000) % 11
    query = torch.nn.functional.normalize(torch.randn(40, 64, generator=gen), dim=1)
    for k in (1, 50, 700):
        pred, scores = _run(query.numpy(), bank.numpy(), labels.numpy(), k, 0.2, 11)
        want, _, _ = K.knn_scores(query.numpy(), bank.numpy(), labels.numpy(), k, 0.2, 11)
        err = np.abs(scores - want).max(axis=1) / want.max(axis=1)
        assert err.max() <= 1e-3, (k, err.max())


def test_errors():
    from medical_image_segmentation_b200 import KNNOnlineEvaluator
    q, b, l = torch.randn(4, 64).cuda(), torch.randn(10, 64).cuda(), torch.zeros(10, dtype=torch.long).cuda()
    with pytest.raises(RuntimeError):                 # k larger than the bank: torch.topk raises as well
        KNNOnlineEvaluator(k=11, num_classes=3).predict(q, b, l)
    with pytest.raises(RuntimeError):                 # no CPU path
        KNNOnlineEvaluator(k=2, num_classes=3).predict(q.cpu(), b.cpu(), l.cpu())
    with pytest.raises(ValueError):
        KNNOnlineEvaluator(k=2, num_classes=3).predict(q, b[:, :32], l)
    with pytest.raises(NotImplementedError):          # D not a multiple of 32
        KNNOnlineEvaluator(k=2, num_classes=3).predict(q[:, :48].contiguous(), b[:, :48].contiguous(), l)
