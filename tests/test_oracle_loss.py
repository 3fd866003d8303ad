"""Loss oracle checks.  BYOL loss is pinned to the reference (byol_pytorch.py:181-198) through
tests/golden/byol_loss.npz; NT-Xent is PARITY UNPINNED (absent from the reference) and
is checked against torch autograd in fp64."""
import os

import numpy as np
import torch

from oracle import loss_oracle as L
from tests import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_byol_loss_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "byol_loss.npz"))
    for i in range(3):
        got = L.byol_cosine_loss(torch.from_numpy(g[f"preds_{i}"]), torch.from_numpy(g[f"targets_{i}"]))
        assert abs(float(got) - float(g[f"loss_{i}"])) <= 1e-6


def test_ntxent_closed_form_matches_autograd():
    for clustered in (False, True):
        z1, z2 = synth.embeddings(48, 32, seed=1, clustered=clustered, dtype=torch.float64)
        a = z1.clone().requires_grad_(True)
        b = z2.clone().requires_grad_(True)
        loss = L.ntxent_loss(a, b, 0.1)
        loss.backward()
        l2, lse, da, db = L.ntxent_closed_form(z1.numpy(), z2.numpy(), 0.1)
        assert abs(float(loss) - l2) < 1e-12
        assert np.abs(a.grad.numpy() - da).max() < 1e-14
        assert np.abs(b.grad.numpy() - db).max() < 1e-14


def test_rank_sharded_convention_equals_global():
    """A.5: rank-major layout gives the same global loss; returned grads are W * dL_global/dz."""
    W, B, D = 4, 6, 16
    g = torch.Generator().manual_seed(3)
    z_locals = [torch.randn(2 * B, D, generator=g, dtype=torch.float64) for _ in range(W)]
    losses, grads = L.ntxent_rank_sharded(z_locals, 0.1)
    z1 = torch.cat([z[:B] for z in z_locals])
    z2 = torch.cat([z[B:] for z in z_locals])
    l_glob, _, d1, d2 = L.ntxent_closed_form(z1.numpy(), z2.numpy(), 0.1)
    assert abs(np.mean(losses) - l_glob) < 1e-12
    for r in range(W):
        assert np.abs(grads[r][:B].numpy() - W * d1[r * B:(r + 1) * B]).max() < 1e-13
        assert np.abs(grads[r][B:].numpy() - W * d2[r * B:(r + 1) * B]).max() < 1e-13


def test_two_independent_ntxent_forms_and_the_committed_vectors_agree():
    """tests/golden/ntxent.npz was produced by the lightning-bolts exp-sum form (oracle/make_ntxent_golden.py); the
    cross-entropy form the kernels were specified against must reproduce it (loss, lse, gradients), and the bolts form
    must reproduce itself from the stored inputs.  (eps = 1e-6 against exp-sums >= 1e2: agreement ~1e-9.)"""
    g = np.load(os.path.join(GOLD, "ntxent.npz"))
    for tag in "abcd":
        z1, z2, T = g[f"{tag}_z1"], g[f"{tag}_z2"], float(g[f"{tag}_T"])
        loss, lse, d1, d2 = L.ntxent_closed_form(z1, z2, T)
        assert abs(loss - float(g[f"{tag}_loss"])) <= 1e-7 * abs(loss)
        assert np.abs(lse - g[f"{tag}_lse"]).max() <= 1e-10
        ref = np.concatenate([g[f"{tag}_dz1"], g[f"{tag}_dz2"]])
        got = np.concatenate([d1, d2])
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-6
        again = L.ntxent_loss_bolts(torch.from_numpy(z1).double(), torch.from_numpy(z2).double(), T)
        assert abs(float(again) - float(g[f"{tag}_loss"])) <= 1e-13
    zl = [torch.from_numpy(z) for z in g["w4_z"]]
    losses, grads = L.ntxent_rank_sharded(zl, float(g["w4_T"]))
    assert np.abs(np.array(losses) - g["w4_loss"]).max() <= 1e-7
    for r in range(4):
        assert np.linalg.norm(grads[r].numpy() - g["w4_dz"][r]) / np.linalg.norm(g["w4_dz"][r]) <= 1e-6
