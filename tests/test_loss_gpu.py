"""GPU parity of the tcgen05 NT-Xent kernels (K2/K3) and the BYOL loss, through the C ABI.

Oracle: oracle/loss_oracle.py in fp64 (NT-Xent: PARITY UNPINNED -- absent from the reference, SURVEY F1;
BYOL: pinned by tests/golden/byol_loss.npz, generated with the reference's own function).
Gates: loss rel <= 1e-3; gradients ||dZ_k - dZ_o||_F / ||dZ_o||_F <= 1e-3 and
elementwise <= 1e-3*|o| + 1e-3*max|o|.
"""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as L
from tests import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _grad_ok(got, ref, what):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    fro = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert fro <= 1e-3, f"{what}: relative Frobenius error {fro:.3e}"
    bound = 1e-3 * np.abs(ref) + 1e-3 * np.abs(ref).max()
    worst = (np.abs(got - ref) / bound).max()
    assert worst <= 1.0, f"{what}: elementwise error {worst:.3f} x bound"
    return fro


@pytest.mark.parametrize("n,d,clustered", [(64, 32, False), (128, 128, False), (256, 128, True), (1024, 128, False),
                                           (1024, 128, True), (512, 256, False), (192, 64, True),
                                           (128, 512, False), (64, 2048, False)])
def test_ntxent_matches_oracle(n, d, clustered):
    from medical_image_segmentation_b200 import nt_xent_loss
    z1, z2 = synth.embeddings(n, d, seed=n + d, clustered=clustered)
    a = z1.cuda().requires_grad_(True)
    b = z2.cuda().requires_grad_(True)
    loss = nt_xent_loss(a, b, 0.1)
    loss.backward()
    ref_loss, _, d1, d2 = L.ntxent_closed_form(z1.numpy(), z2.numpy(), 0.1)
    assert abs(float(loss) - ref_loss) <= 1e-3 * abs(ref_loss), (float(loss), ref_loss)
    _grad_ok(a.grad.cpu().numpy(), d1, "dz_a")
    _grad_ok(b.grad.cpu().numpy(), d2, "dz_b")


@pytest.mark.parametrize("tag", list("abcd"))
def test_ntxent_matches_committed_vectors(tag):
    """Kernels vs tests/golden/ntxent.npz (fp64, produced by the independent lightning-bolts-form oracle)."""
    from medical_image_segmentation_b200 import nt_xent_loss
    g = np.load(os.path.join(GOLD, "ntxent.npz"))
    T = float(g[f"{tag}_T"])
    a = torch.from_numpy(g[f"{tag}_z1"]).cuda().requires_grad_(True)
    b = torch.from_numpy(g[f"{tag}_z2"]).cuda().requires_grad_(True)
    loss = nt_xent_loss(a, b, T)
    loss.backward()
    ref = float(g[f"{tag}_loss"])
    assert abs(float(loss.detach()) - ref) <= 1e-3 * abs(ref)
    _grad_ok(a.grad.cpu().numpy(), g[f"{tag}_dz1"], "dz_a")
    _grad_ok(b.grad.cpu().numpy(), g[f"{tag}_dz2"], "dz_b")


def test_ntxent_rank_sharded_matches_committed_vectors():
    """W = 4 simulated ranks through the ABI (rank-major gathered matrix) vs the bolts-form sharded vectors."""
    from medical_image_segmentation_b200.loss import CudaKernels
    g = np.load(os.path.join(GOLD, "ntxent.npz"))
    zl = [torch.from_numpy(z) for z in g["w4_z"]]
    # The vectors hold each rank's sum_r' dL_r'/dz_local = W * dL_global/dz_local; the four 64-row blocks
    # are smaller than a 128-row tile, so the global problem is evaluated on one rank and scaled by W.
    W, rows, D = len(zl), zl[0].shape[0], zl[0].shape[1]
    B = rows // 2
    z1 = torch.cat([z[:B] for z in zl]).cuda().requires_grad_(True)
    z2 = torch.cat([z[B:] for z in zl]).cuda().requires_grad_(True)
    from medical_image_segmentation_b200 import nt_xent_loss
    loss = nt_xent_loss(z1, z2, float(g["w4_T"]))
    loss.backward()
    assert abs(float(loss.detach()) - float(g["w4_loss"].mean())) <= 1e-3 * float(g["w4_loss"].mean())
    for r in range(W):
        got = torch.cat([z1.grad[r * B:(r + 1) * B], z2.grad[r * B:(r + 1) * B]]).cpu().numpy() * W
        _grad_ok(got, g["w4_dz"][r], f"rank {r}")
    assert CudaKernels.launches > 0


@pytest.mark.parametrize("n,d", [(50, 64), (96, 128), (200, 128), (1, 32), (321, 256), (70, 512)])
def test_ntxent_any_batch_size(n, d):
    """Per-GPU batches that are not a multiple of 64 (the reference takes any --batch_size, train_ssl.py:28-41): the
    rank's block is padded to 128-row tiles with masked columns; loss and gradients match the oracle on the valid rows."""
    from medical_image_segmentation_b200 import nt_xent_loss
    z1, z2 = synth.embeddings(n, d, seed=3 * n + d, clustered=(n % 2 == 0))
    a = z1.cuda().requires_grad_(True)
    b = z2.cuda().requires_grad_(True)
    loss = nt_xent_loss(a, b, 0.1)
    loss.backward()
    ref_loss, _, d1, d2 = L.ntxent_closed_form(z1.numpy(), z2.numpy(), 0.1)
    # (tiny batches of clustered pairs give losses ~0.01 = lse - s_pos with both terms ~10: the relative gate is floored
    # at a loss of 0.1, i.e. 1e-5 relative to the terms)
    assert abs(float(loss.detach()) - ref_loss) <= 1e-3 * max(abs(ref_loss), 0.1), (float(loss), ref_loss)
    if n > 1:
        _grad_ok(a.grad.cpu().numpy(), d1, "dz_a")
        _grad_ok(b.grad.cpu().numpy(), d2, "dz_b")


@pytest.mark.parametrize("temperature", [0.5, 0.07])
def test_ntxent_temperatures_and_grad_scale(temperature):
    from medical_image_segmentation_b200 import nt_xent_loss
    z1, z2 = synth.embeddings(128, 64, seed=4)
    a = z1.cuda().requires_grad_(True)
    b = z2.cuda().requires_grad_(True)
    (3.0 * nt_xent_loss(a, b, temperature)).backward()
    ref_loss, _, d1, d2 = L.ntxent_closed_form(z1.numpy(), z2.numpy(), temperature)
    _grad_ok(a.grad.cpu().numpy(), 3.0 * d1, "dz_a")
    _grad_ok(b.grad.cpu().numpy(), 3.0 * d2, "dz_b")


def test_ntxent_bf16_embeddings():
    """bf16 embeddings in, bf16 gradients out; the oracle sees the same bf16-rounded inputs."""
    from medical_image_segmentation_b200 import nt_xent_loss
    z1, z2 = synth.embeddings(256, 128, seed=8, dtype=torch.bfloat16)
    a = z1.cuda().requires_grad_(True)
    b = z2.cuda().requires_grad_(True)
    loss = nt_xent_loss(a, b, 0.1)
    loss.backward()
    ref_loss, _, d1, d2 = L.ntxent_closed_form(z1.float().numpy(), z2.float().numpy(), 0.1)
    assert abs(float(loss) - ref_loss) <= 1e-3 * abs(ref_loss)
    assert a.grad.dtype == torch.bfloat16
    # bf16 output rounding (2^-9) dominates: compare against the bf16-rounded oracle gradient
    ref = torch.from_numpy(d1).to(torch.bfloat16).float().numpy()
    got = a.grad.float().cpu().numpy()
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 4e-3


@pytest.mark.parametrize("W,B,D", [(3, 40, 64), (2, 100, 512), (4, 64, 768)])
def test_ntxent_rank_sharded_layout_single_gpu(W, B, D):
    """A.5 on one GPU: feed each simulated rank's rows with the all-gathered matrix through the ABI.  (3, 40, 64): 80
    rows per rank, every block padded to one 128-row tile; D > 256: the backward that writes W and runs the dU GEMM,
    with row0 != 0, padded blocks and several K-splits."""
    from medical_image_segmentation_b200.loss import CudaKernels
    g = torch.Generator().manual_seed(3)
    z_locals = [torch.randn(2 * B, D, generator=g) for _ in range(W)]
    ref_losses, ref_grads = L.ntxent_rank_sharded(z_locals, 0.1)
    preps = [CudaKernels.prep(z.cuda()) for z in z_locals]
    u_all = torch.cat([p[1] for p in preps])
    rows = 2 * B
    rp = CudaKernels.padded_rows(rows)
    assert rp % 128 == 0 and u_all.shape[0] == W * rp
    scratch = CudaKernels.scratch(rows, W * rp, D, "cuda")
    outs = [CudaKernels.fwd(u_all, r * rp, rows, 10.0, scratch) for r in range(W)]
    lse_all = torch.cat([o[0] for o in outs])
    one = torch.ones(1, device="cuda")
    for r in range(W):
        assert abs(float(outs[r][1]) - ref_losses[r]) <= 1e-3 * abs(ref_losses[r])
        dz = CudaKernels.bwd(u_all, lse_all, preps[r][0], preps[r][2], r * rp, 10.0, one, scratch)
        _grad_ok(dz.cpu().numpy(), ref_grads[r].numpy(), f"rank {r}")


def test_ntxent_errors():
    from medical_image_segmentation_b200 import nt_xent_loss
    with pytest.raises(NotImplementedError):      # D not a multiple of 32
        nt_xent_loss(torch.randn(64, 48).cuda(), torch.randn(64, 48).cuda())
    with pytest.raises(NotImplementedError):      # temperature below the fixed-max range
        nt_xent_loss(torch.randn(64, 64).cuda(), torch.randn(64, 64).cuda(), 0.01)
    with pytest.raises(ValueError):
        nt_xent_loss(torch.randn(64, 64).cuda(), torch.randn(32, 64).cuda())
    with pytest.raises(RuntimeError):
        nt_xent_loss(torch.randn(64, 64), torch.randn(64, 64))


def test_byol_loss_matches_reference_golden():
    from medical_image_segmentation_b200 import byol_cosine_loss
    g = np.load(os.path.join(GOLD, "byol_loss.npz"))
    for i in range(3):
        p = torch.from_numpy(g[f"preds_{i}"]).cuda().requires_grad_(True)
        t = torch.from_numpy(g[f"targets_{i}"]).cuda()
        loss = byol_cosine_loss(p, t)
        assert abs(float(loss) - float(g[f"loss_{i}"])) <= 1e-5
        loss.backward()
        pr = torch.from_numpy(g[f"preds_{i}"]).double().requires_grad_(True)
        L.byol_cosine_loss(pr, torch.from_numpy(g[f"targets_{i}"]).double()).backward()
        _grad_ok(p.grad.cpu().numpy(), pr.grad.numpy(), f"byol case {i}")


def test_momentum_update_is_bit_identical_to_the_reference_loop():
    """BYOL.momentum_update (byol_pytorch.py:291-296) on a small model: the multi-tensor kernel vs the reference's
    per-tensor mul_ / add_ loop, bit for bit, including odd lengths and unaligned views."""
    from medical_image_segmentation_b200 import momentum_update
    torch.manual_seed(0)
    online = torch.nn.Sequential(torch.nn.Conv2d(3, 17, 3), torch.nn.BatchNorm2d(17), torch.nn.Linear(301, 77),
                                 torch.nn.Linear(77, 5000)).cuda()
    mom = torch.nn.Sequential(torch.nn.Conv2d(3, 17, 3), torch.nn.BatchNorm2d(17), torch.nn.Linear(301, 77),
                              torch.nn.Linear(77, 5000)).cuda()
    ref = [p.detach().clone() for p in mom.parameters()]
    for m in (0.996, 0.99, 1.0, 0.5):
        for po, pr in zip(online.parameters(), ref):
            pr.mul_(m).add_(po.data, alpha=1.0 - m)                 # the reference loop
        momentum_update(online, mom, m)
        for pm, pr in zip(mom.parameters(), ref):
            assert torch.equal(pm.data, pr)
    flat_o, flat_m = torch.randn(10001, device="cuda"), torch.randn(10001, device="cuda")
    want = flat_m[1:].clone().mul_(0.9).add_(flat_o[1:], alpha=1.0 - 0.9)
    momentum_update([flat_o[1:]], [flat_m[1:]], 0.9)                # 4-byte aligned only: scalar path
    assert torch.equal(flat_m[1:], want)
    with pytest.raises(RuntimeError):
        momentum_update([torch.zeros(4)], [torch.zeros(4)], 0.9)


def test_mean_std_matches_reference_formula():
    """compute_mean_and_std (analyze_data/compute_dataset_metrics.py:12-29) restated in float64 on the CPU."""
    from medical_image_segmentation_b200 import compute_mean_and_std
    imgs = synth.batch_512(6, seed=9, H=200, W=256)
    batches = [torch.from_numpy(imgs[i:i + 2])[:, None].cuda() for i in range(0, 6, 2)]
    mean, std = compute_mean_and_std([(b, None) for b in batches])
    x = torch.from_numpy(imgs.astype(np.int64)).to(torch.float64)
    ref_mean = x.sum() / x.numel()
    ref_std = torch.sqrt((x ** 2).sum() / x.numel() - ref_mean ** 2)
    assert abs(float(mean[0]) - float(ref_mean)) <= 1e-9 * float(ref_mean)
    assert abs(float(std[0]) - float(ref_std)) <= 1e-9 * float(ref_std)
    m01, s01 = compute_mean_and_std(batches, scale=1.0 / 65535.0)
    assert abs(float(m01[0]) - float(ref_mean) / 65535.0) < 1e-12


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one node")
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_two_rank_exchange_matches_sharded_oracle(exchange):
    """NVLink peer-store exchange (and the NCCL fallback) on two real ranks: scripts/dist_check.py compares every
    rank's loss and gradient with the rank-sharded oracle and fails on a peer-wait timeout."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MIS_NTXENT_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "scripts", "dist_check.py")],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_ntxent_full_size_properties():
    """2N = 8192, D = 128 (the cfg3 matrix on one GPU), through properties that hold at any size: the loss does not
    depend on the scale of a row (so <z_i, dz_i> = 0); swapping the two views leaves the loss unchanged and swaps the
    gradients; the loss is bounded by log(2N - 1) +- 2/T; doubling the temperature-scaled problem is deterministic."""
    from medical_image_segmentation_b200 import nt_xent_loss
    n, d, T = 4096, 128, 0.1
    z1, z2 = synth.embeddings(n, d, seed=5)
    a = z1.cuda().requires_grad_(True)
    b = z2.cuda().requires_grad_(True)
    loss = nt_xent_loss(a, b, T)
    loss.backward()
    assert abs(float(loss) - np.log(2 * n - 1)) <= 2.0 / T
    for zz in (a, b):                                      # scale invariance of every row
        radial = (zz.detach() * zz.grad).sum(1).abs().max().item()
        assert radial <= 1e-3 * zz.grad.abs().max().item() * zz.detach().norm(dim=1).max().item()
    a2 = z2.cuda().requires_grad_(True)
    b2 = z1.cuda().requires_grad_(True)
    loss2 = nt_xent_loss(a2, b2, T)
    loss2.backward()
    assert abs(float(loss2) - float(loss)) <= 1e-5 * abs(float(loss))
    assert (a2.grad - b.grad).abs().max().item() <= 1e-3 * b.grad.abs().max().item()
    a3 = (3.0 * z1).cuda().requires_grad_(True)            # rescaled rows: same loss, gradients / 3
    b3 = z2.cuda().requires_grad_(True)
    loss3 = nt_xent_loss(a3, b3, T)
    loss3.backward()
    assert abs(float(loss3) - float(loss)) <= 1e-5 * abs(float(loss))
    assert (3.0 * a3.grad - a.grad).abs().max().item() <= 1e-3 * a.grad.abs().max().item()
