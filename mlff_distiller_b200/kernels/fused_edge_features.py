"""Drop-in for the reference's ``kernels/fused_edge_features.py``: edge vectors, distances and unit
vectors in one launch (``mlffd_edge_features``, csrc/edge_features.cuh)."""
from __future__ import annotations

from typing import Tuple

import torch

from .. import _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise _lib.MlffdError(rc, f"{what} failed")


def fused_edge_features_triton(positions: torch.Tensor, edge_index: torch.Tensor, eps: float = 1e-8,
                               *, eps_placement: str = "triton"
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``(edge_vectors [E,3], distances [E], normalized_vectors [E,3])`` for ``edge_index [2,E]``
    (row 0 = src, row 1 = dst), ``edge_vec = pos[src] - pos[dst]``: the signature and semantics of
    kernels/fused_edge_features.py:99-162.

    ``eps_placement='triton'`` (default) reproduces the Triton kernel, ``d = sqrt(|r|^2 + eps)``,
    ``u = r / d`` (fused_edge_features.py:77-82); ``'model'`` reproduces the model's own ops,
    ``d = |r|``, ``u = r / (d + eps)`` (student_model.py:712-715) -- what the fused product path uses.
    """
    if eps_placement not in ("triton", "model"):
        raise ValueError("eps_placement must be 'triton' or 'model'")
    if positions.device.type != "cuda":
        raise RuntimeError("fused_edge_features_triton runs on CUDA only (no CPU fallback)")
    if positions.dtype != torch.float32:
        raise TypeError("positions must be float32")
    lib = _lib.load()
    pos = positions.contiguous()
    ei = edge_index.to(device=pos.device, dtype=torch.int64).contiguous()
    e = int(ei.shape[1])
    vec = torch.empty((e, 3), dtype=torch.float32, device=pos.device)
    dist = torch.empty(e, dtype=torch.float32, device=pos.device)
    unit = torch.empty((e, 3), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        _check(lib.mlffd_edge_features(pos.data_ptr(), ei.data_ptr(), e, float(eps),
                                       1 if eps_placement == "triton" else 0, vec.data_ptr(), dist.data_ptr(),
                                       unit.data_ptr(), torch.cuda.current_stream(pos.device).cuda_stream),
               "mlffd_edge_features")
    return vec, dist, unit
