"""Stand-in for loss.CudaKernels used ONLY by the CPU gloo tests of the multi-rank plumbing.

It implements the same three steps (prep / fwd / bwd, SURVEY A.5 option L) with plain torch fp64
math so that the host logic around the kernels -- rank-major all-gather, row offsets, lse gather,
gradient convention -- can be exercised without a GPU.  Test infrastructure, never shipped."""
import torch


class TorchKernels:
    @staticmethod
    def padded_rows(rows):
        return rows          # the stand-in needs no 128-row tiles

    @staticmethod
    def prep(z):
        z = z.contiguous()
        zd = z.double()
        rinv = 1.0 / zd.norm(dim=1).clamp_min(1e-12)
        return z, zd * rinv[:, None], rinv

    @staticmethod
    def scratch(rows, cols, D, device):
        return torch.empty(0)

    @staticmethod
    def fwd(u_all, row0, rows, inv_T, scratch):
        s = (u_all[row0:row0 + rows] @ u_all.T) * inv_T
        idx = torch.arange(rows)
        s[idx, row0 + idx] = float("-inf")
        lse = torch.logsumexp(s, dim=1)
        pos = s[idx, row0 + (idx + rows // 2) % rows]
        return lse, (lse - pos).mean().reshape(1)

    @staticmethod
    def bwd(u_all, lse_all, z, rinv, row0, inv_T, grad_out, scratch):
        rows = z.shape[0]
        u = u_all[row0:row0 + rows]
        s = (u @ u_all.T) * inv_T
        idx = torch.arange(rows)
        s[idx, row0 + idx] = float("-inf")
        w = torch.exp(s - lse_all[row0:row0 + rows, None]) + torch.exp(s - lse_all[None, :])
        w[idx, row0 + (idx + rows // 2) % rows] -= 2.0
        du = (w @ u_all) * (inv_T / rows) * grad_out.double()
        dz = (du - u * (u * du).sum(1, keepdim=True)) * rinv[:, None]
        return dz.to(z.dtype)
