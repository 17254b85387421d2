"""Device engine: one libmlffd context per GPU, fed with torch tensors as raw device pointers.

PyTorch is plumbing here (device memory, streams); all arithmetic of the path runs in the
hand-written CUDA kernels behind the C ABI (include/mlffd.h).
"""
from __future__ import annotations

import ctypes
from typing import Mapping, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .checkpoint import ModelConfig, pack_weights


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _DeviceArray:
    """Zero-copy view of library-owned device memory through __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr,
                                         "data": (ptr, False), "version": 2}


class Engine:
    """Owns a ``mlffd_ctx`` on one CUDA device.

    All ``*_async`` methods only enqueue work on the current torch stream of the device.
    """

    def __init__(self, state: Mapping[str, np.ndarray], cfg: ModelConfig, device="cuda",
                 precision: str = "fp32", filter_mode: str = "spline"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the B200 energy+force path needs a CUDA device (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available: the B200 energy+force path has no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        if filter_mode not in _lib.FILTER_MODES:
            raise ValueError(f"filter_mode must be one of {sorted(_lib.FILTER_MODES)}")
        self.lib = _lib.load()
        self.cfg = cfg
        self.precision = precision
        self.filter_mode = filter_mode
        blob = np.ascontiguousarray(pack_weights(state, cfg), dtype=np.float32)
        c = _lib.MlffdConfig(cfg.hidden_dim, cfg.num_rbf, cfg.num_interactions, cfg.max_z,
                             float(cfg.cutoff), _lib.PRECISIONS[precision],
                             _lib.FILTER_MODES[filter_mode])
        handle = ctypes.c_void_p()
        rc = self.lib.mlffd_model_create(ctypes.byref(handle), self.device.index, ctypes.byref(c),
                                         blob.ctypes.data_as(ctypes.c_void_p), blob.size)
        if rc != 0:
            msg = self.lib.mlffd_last_error(None).decode()
            raise _lib.MlffdError(rc, msg)
        self._ctx = handle
        self.profiling = False
        self.dense_fallback = False
        self.skin = 0.0
        self.saturation_reruns = 0   # calls repeated on the FP32 kernels because a tensor-core operand saturated
        self.cap_atoms = self.cap_edges = self.cap_structs = 0

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.mlffd_model_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise _lib.MlffdError(rc, self.lib.mlffd_last_error(self._ctx).decode())

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- workspace --------------------------------------------------------------------------
    def reserve(self, max_atoms: int, max_edges: int, max_structs: int):
        max_atoms = max(int(max_atoms), self.cap_atoms)
        max_edges = max(int(max_edges), self.cap_edges)
        max_structs = max(int(max_structs), self.cap_structs)
        if (max_atoms, max_edges, max_structs) == (self.cap_atoms, self.cap_edges, self.cap_structs):
            return
        self._check(self.lib.mlffd_workspace_reserve(self._ctx, max_atoms, max_edges, max_structs))
        self.cap_atoms, self.cap_edges, self.cap_structs = max_atoms, max_edges, max_structs

    def ensure(self, n_atoms: int, n_structs: int, edges_per_atom: int = 40):
        """Grow the workspace for a request; the edge capacity is a guess that
        :meth:`status` verifies after the step (overflow -> reserve exact and rerun)."""
        if n_atoms > self.cap_atoms or n_structs > self.cap_structs or self.cap_edges == 0:
            want_e = max(self.cap_edges, min(n_atoms * edges_per_atom, n_atoms * max(n_atoms - 1, 1)), 64)
            self.reserve(max(n_atoms, self.cap_atoms), want_e, max(n_structs, self.cap_structs))

    # -- hot path ---------------------------------------------------------------------------
    def energy_forces_async(self, z: torch.Tensor, pos: torch.Tensor, offsets: torch.Tensor,
                            n_structs: int, energy: torch.Tensor, forces: Optional[torch.Tensor],
                            cells: Optional[torch.Tensor] = None, pbc: Optional[torch.Tensor] = None):
        """Enqueue one E(+F) evaluation.  z int32 [N], pos f32 [N,3], offsets int32 [B+1],
        energy f32 [B] out, forces f32 [N,3] out or None, cells f32 [B,18], pbc uint8 [B,3]."""
        self._check(self.lib.mlffd_energy_forces(
            self._ctx, _ptr(z), _ptr(pos), _ptr(offsets), int(n_structs), int(pos.shape[0]),
            _ptr(cells), _ptr(pbc), _ptr(energy), _ptr(forces), self._stream()))

    def virial_async(self, offsets: torch.Tensor, n_structs: int, virial: torch.Tensor):
        """dE_b / d strain [B,3,3] of the last energy_forces_async call (same stream, forces requested)."""
        self._check(self.lib.mlffd_virial(self._ctx, _ptr(offsets), int(n_structs), _ptr(virial), self._stream()))

    def neighbor_list_async(self, pos: torch.Tensor, offsets: torch.Tensor, n_structs: int,
                            cells: Optional[torch.Tensor] = None, pbc: Optional[torch.Tensor] = None):
        self._check(self.lib.mlffd_neighbor_list(
            self._ctx, _ptr(pos), _ptr(offsets), int(n_structs), int(pos.shape[0]), _ptr(cells),
            _ptr(pbc), self._stream()))

    def set_skin(self, skin: float):
        """Verlet-skin width in Angstrom (0 = exact rebuild every call).  Frees the workspace."""
        self._check(self.lib.mlffd_set_skin(self._ctx, float(skin)))
        self.skin = float(skin)
        self.cap_atoms = self.cap_edges = self.cap_structs = 0

    def set_dense_fallback(self, enable: bool):
        """Run the dense layers on the FP32 FFMA kernels (True) or as the precision says (False).
        Used to repeat a step whose tensor-core operands left the FP16 range (status.tc_saturated)."""
        self._check(self.lib.mlffd_set_dense_fallback(self._ctx, 1 if enable else 0))
        self.dense_fallback = bool(enable)

    def status(self) -> _lib.MlffdStatus:
        """Synchronises the last call's stream and returns its counters."""
        s = _lib.MlffdStatus()
        self._check(self.lib.mlffd_get_status(self._ctx, ctypes.byref(s)))
        return s

    def status_async(self, out: torch.Tensor):
        """Enqueue a copy of the step's six status words (num_edges, num_pairs, overflow,
        max_degree, overflow_events, tc_saturated) into ``out`` (int32[6], pinned host or device)
        behind the step just enqueued on the current stream.  Never synchronises."""
        self._check(self.lib.mlffd_status_async(self._ctx, out.data_ptr(), self._stream()))

    def profile_enable(self, enable: bool = True):
        """Reset launch counters; with ``enable`` also time every stage with CUDA events."""
        self._check(self.lib.mlffd_profile_enable(self._ctx, 1 if enable else 0))
        self.profiling = bool(enable)   # per-kernel events cannot be recorded inside a graph capture

    def profile_read(self) -> dict:
        """{'launches': int, 'stages': {name: {'ms': float, 'launches': int}}} since enable."""
        prof = _lib.MlffdProfile()
        self._check(self.lib.mlffd_profile_read(self._ctx, ctypes.byref(prof)))
        stages = {}
        for i in range(_lib.NUM_STAGES):
            name = self.lib.mlffd_stage_name(i).decode()
            stages[name] = {"ms": float(prof.stage_ms[i]), "launches": int(prof.stage_launches[i])}
        return {"launches": int(prof.launches), "stages": stages}

    def export_edges(self) -> torch.Tensor:
        """edge_index [2,E] int64 of the last neighbour build, in the reference's order."""
        s = self.status()
        if s.overflow:
            raise _lib.MlffdError(_lib.MLFFD_ECAPACITY, "edge capacity exceeded")
        e = int(s.num_edges)
        out = torch.empty((2, max(e, 1)), dtype=torch.int64, device=self.device)
        n = ctypes.c_int64()
        self._check(self.lib.mlffd_export_edges(self._ctx, out.data_ptr(), out.shape[1],
                                                ctypes.byref(n), self._stream()))
        return out[:, :e]

    def filter_table(self, layer: int, dist: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        h = self.cfg.hidden_dim
        dist = dist.to(device=self.device, dtype=torch.float32).contiguous()
        f = torch.empty((dist.numel(), 3 * h), dtype=torch.float32, device=self.device)
        df = torch.empty_like(f)
        self._check(self.lib.mlffd_filter_table(self._ctx, int(layer), dist.data_ptr(),
                                                dist.numel(), f.data_ptr(), df.data_ptr(),
                                                self._stream()))
        return f, df

    def filter_spline(self, layer: int, dist: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Value and d-derivative of the per-model filter spline (what the spline-mode message
        kernels evaluate in shared memory) at ``dist``: two [P, 3H] tensors."""
        h = self.cfg.hidden_dim
        dist = dist.to(device=self.device, dtype=torch.float32).contiguous()
        f = torch.empty((dist.numel(), 3 * h), dtype=torch.float32, device=self.device)
        df = torch.empty_like(f)
        self._check(self.lib.mlffd_filter_spline(self._ctx, int(layer), dist.data_ptr(),
                                                 dist.numel(), f.data_ptr(), df.data_ptr(),
                                                 self._stream()))
        return f, df

    def debug_buffer(self, name: str, layer: int = 0) -> torch.Tensor:
        """Copy of an internal buffer of the last call (tests only)."""
        ptr, cnt, es = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
        self._check(self.lib.mlffd_debug_buffer(self._ctx, name.encode(), int(layer),
                                                ctypes.byref(ptr), ctypes.byref(cnt), ctypes.byref(es)))
        n = int(cnt.value)
        is_int = name in ("rowptr", "col", "rev", "pair", "edge_dst")
        per = es.value // 4
        if n == 0:
            out = torch.empty(0, dtype=torch.int32 if is_int else torch.float32, device=self.device)
        else:
            view = _DeviceArray(ptr.value, n * per, "<i4" if is_int else "<f4")
            out = torch.as_tensor(view, device=self.device).clone()
        return out.view(n, per) if per > 1 else out
