"""Drop-in for the reference's ``kernels/fused_rbf_cutoff.py``: Gaussian RBF expansion times the cosine
cutoff in one launch (``mlffd_rbf_cutoff``, csrc/edge_features.cuh)."""
from __future__ import annotations

import torch

from .. import _lib


def fused_rbf_cutoff_triton(distances: torch.Tensor, centers: torch.Tensor, gamma: float, r_cut: float,
                            block_size: int = 64) -> torch.Tensor:
    """``[E, K]`` values ``exp(-gamma (d - mu_k)^2) * 0.5 (cos(pi d / r_cut) + 1) [d < r_cut]``:
    signature and semantics of kernels/fused_rbf_cutoff.py:90-134 (``block_size`` is accepted and unused)."""
    if distances.device.type != "cuda":
        raise RuntimeError("fused_rbf_cutoff_triton runs on CUDA only (no CPU fallback)")
    if distances.dtype != torch.float32:
        raise TypeError("distances must be float32")
    lib = _lib.load()
    d = distances.contiguous()
    c = centers.to(device=d.device, dtype=torch.float32).contiguous()
    out = torch.empty((d.shape[0], c.shape[0]), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        rc = lib.mlffd_rbf_cutoff(d.data_ptr(), int(d.shape[0]), c.data_ptr(), int(c.shape[0]), float(gamma),
                                  float(r_cut), out.data_ptr(), torch.cuda.current_stream(d.device).cuda_stream)
    if rc != 0:
        raise _lib.MlffdError(rc, "mlffd_rbf_cutoff failed")
    return out


class FusedRBFCutoff(torch.nn.Module):
    """Module wrapper with the constructor, buffers and forward of kernels/fused_rbf_cutoff.py:137-196
    (``centers = linspace(0, cutoff, num_rbf)``, ``widths = cutoff / num_rbf``, ``gamma = 1 / widths[0]^2``)."""

    def __init__(self, num_rbf: int = 20, cutoff: float = 5.0, learnable: bool = False):
        super().__init__()
        self.num_rbf = num_rbf
        self.cutoff = cutoff
        centers = torch.linspace(0, cutoff, num_rbf)
        widths = torch.ones(num_rbf) * (cutoff / num_rbf)
        if learnable:
            self.centers = torch.nn.Parameter(centers)
            self.widths = torch.nn.Parameter(widths)
        else:
            self.register_buffer("centers", centers)
            self.register_buffer("widths", widths)

    def forward(self, distances: torch.Tensor) -> torch.Tensor:
        gamma = 1.0 / (self.widths[0] ** 2)
        return fused_rbf_cutoff_triton(distances, self.centers, gamma.item(), self.cutoff)
