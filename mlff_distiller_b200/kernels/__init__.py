"""Kernel-level API of the reference (``kernels/`` package: fused_edge_features.py, fused_rbf_cutoff.py)
on the hand-written CUDA stage kernels of libmlffd.so (csrc/edge_features.cuh).  Same function names,
argument meaning and return values as the Triton originals, so ``StudentForceFieldOptimized``-style
call sites (student_model_optimized.py:117-140) keep working; there is no Triton and no CPU fallback."""
from .fused_edge_features import fused_edge_features_triton  # noqa: F401
from .fused_rbf_cutoff import FusedRBFCutoff, fused_rbf_cutoff_triton  # noqa: F401
