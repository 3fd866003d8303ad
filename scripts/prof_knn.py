"""Timing of the kNN evaluator at a realistic size: python scripts/prof_knn.py [B] [N] [D] [C] [k]."""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_segmentation_b200 import KNNOnlineEvaluator
B, N, D, Cn, k = (int(a) for a in (sys.argv[1:6] + ["256", "100000", "128", "10", "200"][len(sys.argv) - 1:]))
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=1)
bank = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1)
labels = torch.randint(0, Cn, (N,), device="cuda", generator=g)
ev = KNNOnlineEvaluator(k=k, temperature=0.07, num_classes=Cn)
for _ in range(3):
    pred = ev.predict(q, bank, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    pred = ev.predict(q, bank, labels)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
# the reference's formulation in torch on the same GPU (knn.py:52-70), for scale
def torch_predict():
    sim = q @ bank.T
    w, idx = sim.topk(k=k, dim=-1)
    lab = torch.gather(labels.expand(B, -1), dim=-1, index=idx)
    w = (w / 0.07).exp()
    one_hot = torch.zeros(B * k, Cn, device="cuda").scatter(dim=-1, index=lab.view(-1, 1), value=1.0)
    return torch.sum(one_hot.view(B, -1, Cn) * w.unsqueeze(dim=-1), dim=1).argsort(dim=-1, descending=True)
for _ in range(3):
    ref = torch_predict()
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    ref = torch_predict()
e1.record(); torch.cuda.synchronize()
agree = (ref[:, 0] == pred[:, 0]).float().mean().item()
print(f"kNN predict B={B} N={N} D={D} C={Cn} k={k}: {ms:.3f} ms (torch formulation on the same GPU: {e0.elapsed_time(e1) / 10:.3f} ms); top-1 agreement {agree:.4f}")
