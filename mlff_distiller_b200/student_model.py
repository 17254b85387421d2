"""Host-side mirror of the reference ``StudentForceField`` running on the B200 CUDA path.

Keeps the public surface of src/mlff_distiller/models/student_model.py:532-1176 (paths relative
to /root/reference): constructor arguments, attributes, ``forward`` / ``predict_energy_and_forces``
/ ``forward_with_analytical_forces`` signatures and return shapes, ``num_parameters``,
``save`` / ``load`` (checkpoint dict layout, shape-inferred config, ``model.`` prefix repair) and
an identical ``state_dict`` (35 tensors, same names) so ``load_state_dict`` round-trips with the
reference.  The sub-modules below are parameter CONTAINERS only: no torch op computes on them;
``forward`` hands raw device pointers to libmlffd.so (include/mlffd.h), which evaluates the energy
and the analytical forces in hand-written sm_100a kernels.

``forward`` stays differentiable w.r.t. ``positions`` for callers that follow the reference recipe
``forces = -torch.autograd.grad(energy, positions)`` (predict_energy_and_forces :782-793,
inference/ase_calculator.py:757-763): the autograd node simply returns the analytical forces the
CUDA reverse pass already produced.

Differences that are deliberate and documented: ``pbc_mode='ignore'`` (default)
reproduces the reference, which accepts ``cell``/``pbc`` and ignores them (:694-703);
``pbc_mode='minimum_image'`` enables the periodic neighbour list.
"""
from __future__ import annotations

import logging
import math
from pathlib import Path
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import checkpoint as ckpt
from .checkpoint import ModelConfig

logger = logging.getLogger(__name__)


class _RBFBuffers(nn.Module):
    """Holds ``rbf.centers`` / ``rbf.widths`` exactly as GaussianRBF registers them
    (student_model.py:223-234)."""

    def __init__(self, num_rbf: int, cutoff: float, learnable: bool):
        super().__init__()
        centers = torch.linspace(0, cutoff, num_rbf)
        widths = torch.ones(num_rbf) * (cutoff / num_rbf)
        if learnable:
            self.centers = nn.Parameter(centers)
            self.widths = nn.Parameter(widths)
        else:
            self.register_buffer("centers", centers)
            self.register_buffer("widths", widths)


def _mlp(sizes, n_layers_with_act):
    mods = []
    for i in range(len(sizes) - 1):
        mods.append(nn.Linear(sizes[i], sizes[i + 1]))
        if i < n_layers_with_act:
            mods.append(nn.SiLU())
    return nn.Sequential(*mods)


class _MessageParams(nn.Module):
    def __init__(self, h: int, k: int):
        super().__init__()
        self.rbf_to_scalar = _mlp([k, h, 3 * h], 1)


class _UpdateParams(nn.Module):
    def __init__(self, h: int):
        super().__init__()
        self.update_mlp = _mlp([2 * h, h, 3 * h], 1)
        self.mixing_matrix = nn.Parameter(torch.randn(3, 3) / math.sqrt(3))


class _InteractionParams(nn.Module):
    def __init__(self, h: int, k: int):
        super().__init__()
        self.message = _MessageParams(h, k)
        self.update = _UpdateParams(h)


def check_minimum_image(cell, pbc, cutoff: float, skin: float = 0.0) -> None:
    """The one-image neighbour kernels are exact only if every PERIODIC cell height is >= 2 (r_c + skin);
    a smaller cell would silently lose periodic images.  ``cell`` [3,3] or [B,3,3], ``pbc`` [3] or [B,3]
    (arrays or tensors, any device); raises ``ValueError`` naming the first offending structure and axis.
    Shared by every entry point that accepts a cell: the calculator, ``StudentForceField`` (forward,
    analytical forces, stress), ``radius_graph`` and ``md.DeviceMD``."""
    c = np.asarray(torch.as_tensor(cell).detach().cpu().numpy(), dtype=np.float64).reshape(-1, 3, 3)
    p = np.asarray(torch.as_tensor(pbc).detach().cpu().numpy()).astype(bool).reshape(-1, 3)
    for b in range(max(len(c), len(p))):
        cb, pb = c[min(b, len(c) - 1)], p[min(b, len(p) - 1)]
        if not pb.any():
            continue
        vol = abs(np.linalg.det(cb))
        for k in range(3):
            if not pb[k]:
                continue
            area = np.linalg.norm(np.cross(cb[(k + 1) % 3], cb[(k + 2) % 3]))
            height = vol / area if area > 0 else 0.0
            where = f"axis {k}" if max(len(c), len(p)) == 1 else f"structure {b}, axis {k}"
            if height < 2.0 * cutoff:
                raise ValueError(f"pbc_mode='minimum_image' needs cell heights >= 2*cutoff "
                                 f"({2 * cutoff:.2f} Å); {where} has {height:.3f} Å")
            if height < 2.0 * (cutoff + skin):
                raise ValueError(f"skin={skin} needs periodic cell heights >= 2*(cutoff + skin) "
                                 f"({2 * (cutoff + skin):.2f} Å); {where} has {height:.3f} Å")


def _offsets_from_batch(batch: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """offsets [B+1] int32 from a sorted batch vector (one device sync, like the reference's
    ``batch.max()`` at student_model.py:739-745)."""
    nb = int(batch.max().item()) + 1
    counts = torch.bincount(batch, minlength=nb)
    if bool((batch[1:] < batch[:-1]).any().item()):
        raise ValueError("batch indices must be sorted (atoms of a structure contiguous)")
    offsets = torch.zeros(nb + 1, dtype=torch.int32, device=batch.device)
    offsets[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return offsets, nb


class _EnergyFn(torch.autograd.Function):
    """E(positions) whose backward hands out the analytical forces of the CUDA reverse pass."""

    @staticmethod
    def forward(ctx, positions, model, z, offsets, nb, cells, pbc):
        energy, forces = model._run(z, positions.detach(), offsets, nb, cells, pbc, True)
        ctx.save_for_backward(forces, offsets)
        ctx.nb = nb
        return energy

    @staticmethod
    def backward(ctx, grad_e):
        forces, offsets = ctx.saved_tensors
        counts = (offsets[1:] - offsets[:-1]).to(torch.int64)
        g = torch.repeat_interleave(grad_e.reshape(-1), counts, output_size=forces.shape[0])
        return -forces * g.unsqueeze(1), None, None, None, None, None, None


class StudentForceField(nn.Module):
    """PaiNN student on the hand-written CUDA path (drop-in for the reference class)."""

    def __init__(self, hidden_dim: int = 128, num_interactions: int = 3, num_rbf: int = 20,
                 cutoff: float = 5.0, max_z: int = 118, learnable_rbf: bool = False,
                 use_torch_cluster: bool = True, *, precision: str = "tc",
                 pbc_mode: str = "ignore", filter_mode: str = "spline", skin: float = 0.0):
        super().__init__()
        if pbc_mode not in ("ignore", "minimum_image"):
            raise ValueError("pbc_mode must be 'ignore' or 'minimum_image'")
        self.hidden_dim = hidden_dim
        self.num_interactions = num_interactions
        self.num_rbf = num_rbf
        self.cutoff = cutoff
        self.max_z = max_z
        self.use_torch_cluster = use_torch_cluster  # accepted for compatibility; unused
        self.precision = precision
        self.pbc_mode = pbc_mode
        self.filter_mode = filter_mode   # 'spline' (per-model filter splines in shared memory) | 'table'
        self.skin = float(skin)          # Verlet-skin width in Angstrom (0 = the list is rebuilt exactly every call, as the reference does)
        self.embedding = nn.Embedding(max_z + 1, hidden_dim)
        self.rbf = _RBFBuffers(num_rbf, cutoff, learnable_rbf)
        self.interactions = nn.ModuleList(
            [_InteractionParams(hidden_dim, num_rbf) for _ in range(num_interactions)])
        self.energy_head = _mlp([hidden_dim, hidden_dim // 2, hidden_dim // 4, 1], 2)
        self._initialize_parameters()
        for p in self.parameters():
            p.requires_grad_(False)  # inference path: weights are constants of the kernels
        self._engine = None
        self._engine_key = None
        self._edges_per_atom = 40

    # ---- parameters -----------------------------------------------------------------------
    def _initialize_parameters(self):
        """Same distributions as the reference (student_model.py:618-628): embedding
        U(-sqrt3, sqrt3), Xavier-uniform Linear weights, zero biases."""
        nn.init.uniform_(self.embedding.weight, -math.sqrt(3), math.sqrt(3))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def num_parameters(self) -> int:
        return ckpt.num_parameters(self.config)

    @property
    def config(self) -> ModelConfig:
        return ModelConfig(self.hidden_dim, self.num_interactions, self.num_rbf, float(self.cutoff),
                           self.max_z, False, self.use_torch_cluster)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine_key = None  # weights moved / cast: re-pack on next use
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        state_dict = ckpt.strip_prefix(state_dict)
        last = f"interactions.{self.num_interactions - 1}.update.mixing_matrix"
        if last not in state_dict:  # pruned by the ONNX export; dead for E and F
            state_dict = dict(state_dict)
            state_dict[last] = torch.zeros(3, 3)
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._engine_key = None
        return out

    def refresh_weights(self):
        """Re-upload the weights after an in-place edit of the parameters."""
        self._engine_key = None

    # ---- engine ---------------------------------------------------------------------------
    def _device(self) -> torch.device:
        return self.embedding.weight.device

    def engine(self):
        from .engine import Engine
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError(
                "StudentForceField (B200 path) evaluates on CUDA only; move the model with "
                ".to('cuda') -- there is no CPU fallback")
        if self.embedding.weight.dtype != torch.float32:
            raise RuntimeError("the CUDA path computes in float32; got " +
                               str(self.embedding.weight.dtype))
        key = (dev, self.precision, self.filter_mode)
        if self._engine is None or self._engine_key != key:
            state = {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()}
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(state, self.config, dev, self.precision, self.filter_mode)
            if self.skin > 0.0:
                self._engine.set_skin(self.skin)
            self._engine_key = key
        return self._engine

    def _run(self, z, pos, offsets, nb, cells, pbc, want_forces: bool, max_atoms: int = 0):
        """One evaluation through the C ABI; retries with a larger edge workspace on overflow and on
        the FP32 kernels when a tensor-core operand left the FP16 range (``max_atoms`` is accepted for
        compatibility and unused)."""
        eng = self.engine()
        n = pos.shape[0]
        energy = torch.empty(nb, dtype=torch.float32, device=pos.device)
        forces = torch.empty((n, 3), dtype=torch.float32, device=pos.device) if want_forces else None
        eng.ensure(n, nb, self._edges_per_atom)
        for _ in range(4):
            eng.energy_forces_async(z, pos, offsets, nb, energy, forces, cells, pbc)
            st = eng.status()
            if st.tc_saturated and not st.overflow:
                # |activation| >= 8 125 in a split-FP16 dense layer (dense / unphysical input): this call
                # is repeated with the dense layers on the FP32 FFMA kernels
                eng.set_dense_fallback(True)
                try:
                    eng.energy_forces_async(z, pos, offsets, nb, energy, forces, cells, pbc)
                    st = eng.status()
                finally:
                    eng.set_dense_fallback(False)
                eng.saturation_reruns += 1
            if not st.overflow:
                return energy, forces
            eng.reserve(n, int(st.num_edges * 1.25) + 64, nb)
            self._edges_per_atom = max(self._edges_per_atom, int(st.num_edges * 1.25 / max(n, 1)) + 1)
        raise RuntimeError("edge workspace overflow persisted after growing")

    @staticmethod
    def pack_cells(cell: torch.Tensor, pbc: torch.Tensor, nb: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """[B,18] float32 (cell | inverse, inverse computed in FP64) and [B,3] uint8."""
        c = torch.as_tensor(cell, dtype=torch.float64).reshape(-1, 3, 3).cpu()
        if c.shape[0] == 1 and nb > 1:
            c = c.expand(nb, 3, 3)
        p = torch.as_tensor(pbc).reshape(-1, 3).to(torch.bool).cpu()
        if p.shape[0] == 1 and nb > 1:
            p = p.expand(nb, 3)
        cc = c.clone()
        for b in range(cc.shape[0]):  # non-periodic axes may be zero vectors: make invertible
            for k in range(3):
                if not bool(p[b, k]) and float(cc[b, k].abs().sum()) == 0.0:
                    cc[b, k, k] = 1.0
        inv = torch.linalg.inv(cc)
        packed = torch.cat([c.reshape(-1, 9), inv.reshape(-1, 9)], dim=1).to(torch.float32)
        return packed.contiguous().to(device), p.to(torch.uint8).contiguous().to(device)

    def _prepare(self, atomic_numbers, positions, cell, pbc, batch):
        dev = self._device()
        z = atomic_numbers.to(device=dev, dtype=torch.int32).contiguous()
        pos = positions.to(device=dev)
        if pos.dtype != torch.float32:
            raise RuntimeError("positions must be float32 on the CUDA path")
        n = z.shape[0]
        if batch is None:
            nb = 1
            offsets = torch.tensor([0, n], dtype=torch.int32, device=dev)
        else:
            offsets, nb = _offsets_from_batch(batch.to(dev))
        cells_d = pbc_d = None
        if self.pbc_mode == "minimum_image" and cell is not None and pbc is not None \
                and bool(torch.as_tensor(pbc).any()):
            check_minimum_image(cell, pbc, float(self.cutoff), self.skin)
            cells_d, pbc_d = self.pack_cells(cell, pbc, nb, dev)
        return z, pos, offsets, nb, cells_d, pbc_d

    # ---- reference API --------------------------------------------------------------------
    def forward(self, atomic_numbers: torch.Tensor, positions: torch.Tensor,
                cell: Optional[torch.Tensor] = None, pbc: Optional[torch.Tensor] = None,
                batch: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Total energy: 0-dim tensor for one structure, ``[B]`` for a batch
        (student_model.py:634-757)."""
        z, pos, offsets, nb, cells_d, pbc_d = self._prepare(atomic_numbers, positions, cell, pbc, batch)
        if pos.requires_grad and torch.is_grad_enabled():
            e = _EnergyFn.apply(pos.contiguous(), self, z, offsets, nb, cells_d, pbc_d)
        else:
            e, _ = self._run(z, pos.detach().contiguous(), offsets, nb, cells_d, pbc_d, False)
        return e.reshape(()) if (batch is None or nb == 1) else e

    def predict_energy_and_forces(self, atomic_numbers: torch.Tensor, positions: torch.Tensor,
                                  cell: Optional[torch.Tensor] = None,
                                  pbc: Optional[torch.Tensor] = None
                                  ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(E scalar, F [N,3]) -- student_model.py:759-795, with the analytical reverse pass in
        place of autograd."""
        return self.forward_with_analytical_forces(atomic_numbers, positions, cell, pbc, None)

    def forward_with_analytical_forces(self, atomic_numbers, positions, cell=None, pbc=None,
                                       batch=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(E, F) without autograd -- what student_model.py:797-932 set out to be."""
        z, pos, offsets, nb, cells_d, pbc_d = self._prepare(atomic_numbers, positions, cell, pbc, batch)
        e, f = self._run(z, pos.detach().contiguous(), offsets, nb, cells_d, pbc_d, True)
        return (e.reshape(()) if (batch is None or nb == 1) else e), f

    def energy_and_forces_packed(self, z_i32: torch.Tensor, pos_f32: torch.Tensor,
                                 offsets_i32: torch.Tensor, n_structs: int,
                                 cells: Optional[torch.Tensor] = None,
                                 pbc: Optional[torch.Tensor] = None, max_atoms: int = 0):
        """Batched fast path: inputs already in the C-ABI layout, no host sync besides the
        status check.  ``max_atoms`` = size of the largest structure if the caller knows it.
        Returns (E [B], F [N,3])."""
        return self._run(z_i32, pos_f32, offsets_i32, n_structs, cells, pbc, True, max_atoms)

    def virial_of_last_call(self, offsets_i32: torch.Tensor, n_structs: int) -> torch.Tensor:
        """dE_b / d strain, ``[B,3,3]``, of the evaluation that just ran through
        :meth:`energy_and_forces_packed` / :meth:`forward_with_analytical_forces` (same stream).
        W_ac = sum over directed edges of (dE/dr_e)_a (r_e)_c with x -> (1 + eps) x; stress = W / V."""
        w = torch.empty((n_structs, 3, 3), dtype=torch.float32, device=offsets_i32.device)
        self.engine().virial_async(offsets_i32, n_structs, w)
        return w

    def predict_energy_forces_stress(self, atomic_numbers: torch.Tensor, positions: torch.Tensor,
                                     cell: torch.Tensor, pbc: Optional[torch.Tensor] = None
                                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(E, F [N,3], stress [3,3] in eV/A^3) for one structure with a cell.  The stress is the
        symmetrised strain derivative of the energy divided by the cell volume -- the quantity
        inference/ase_calculator.py:521-588 tries to obtain (and returns zeros for, because the
        reference model never reads its ``cell`` argument)."""
        z, pos, offsets, nb, cells_d, pbc_d = self._prepare(atomic_numbers, positions, cell, pbc, None)
        e, f = self._run(z, pos.detach().contiguous(), offsets, nb, cells_d, pbc_d, True)
        w = self.virial_of_last_call(offsets, nb)[0]
        vol = torch.linalg.det(torch.as_tensor(cell, dtype=torch.float64).reshape(3, 3)).abs()
        if float(vol) <= 0.0:
            raise ValueError("stress needs a cell with non-zero volume")
        stress = (0.5 * (w + w.T).double() / vol).to(torch.float32)
        return e.reshape(()), f, stress

    # ---- checkpoints ----------------------------------------------------------------------
    def save(self, path: Union[str, Path]):
        """Reference inference checkpoint layout (student_model.py:1075-1099)."""
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        torch.save({"model_state_dict": {k: v.detach().cpu() for k, v in self.state_dict().items()},
                    "config": self.config.as_dict(), "num_parameters": self.num_parameters()}, path)
        logger.info("Saved model checkpoint to %s", path)

    @classmethod
    def from_state(cls, state: Dict[str, np.ndarray], cfg: ModelConfig, device: str = "cpu",
                   **kwargs) -> "StudentForceField":
        model = cls(hidden_dim=cfg.hidden_dim, num_interactions=cfg.num_interactions,
                    num_rbf=cfg.num_rbf, cutoff=cfg.cutoff, max_z=cfg.max_z,
                    use_torch_cluster=cfg.use_torch_cluster, **kwargs)
        full = ckpt.complete_state(state, cfg)
        model.load_state_dict({k: torch.from_numpy(np.array(v, dtype=np.float32))
                               for k, v in full.items()})
        model.to(device)
        model.eval()
        return model

    @classmethod
    def load(cls, path: Union[str, Path], device: str = "cpu", **kwargs) -> "StudentForceField":
        """Load any supported checkpoint (student_model.py:1101-1176 + ONNX / npz)."""
        state, cfg, meta = ckpt.load_any(path)
        model = cls.from_state(state, cfg, device, **kwargs)
        logger.info("Loaded model from %s (%s parameters)", path,
                    f"{meta.get('num_parameters', model.num_parameters()):,}")
        return model

    @classmethod
    def from_student_model(cls, student, **kwargs) -> "StudentForceField":
        """Optimized-subclass idiom of the reference (student_model_optimized.py:174-212)."""
        cfg = ModelConfig(student.hidden_dim, student.num_interactions, student.num_rbf,
                          float(student.cutoff), student.max_z)
        state = {k: v.detach().cpu().numpy() for k, v in student.state_dict().items()}
        dev = next(student.parameters()).device
        return cls.from_state(state, cfg, str(dev), **kwargs)


class EnergyOnlyWrapper(nn.Module):
    """``(atomic_numbers, positions) -> energy``: the two-argument signature of the reference's
    TorchScript export (``SimpleWrapper``, scripts/export_to_torchscript.py:77-84) that
    ``StudentForceFieldCalculator(use_jit=True)`` calls before differentiating the energy with
    respect to the positions (inference/ase_calculator.py:319-335).  ``state_dict`` keys carry the
    same ``model.`` prefix as the export; the backward pass is the fused reverse kernels."""

    def __init__(self, base_model: "StudentForceField"):
        super().__init__()
        self.model = base_model

    def forward(self, atomic_numbers: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        return self.model(atomic_numbers, positions, None, None, None)


_GRAPH_ENGINES: Dict[Tuple[int, float], "object"] = {}


def _graph_engine(device: torch.device, cutoff: float):
    """Weight-free CUDA context for neighbour lists only, cached per (device, cutoff): the neighbour
    kernels read nothing but the cutoff, so a minimal all-zero model (H = 32, one layer) carries it."""
    from .engine import Engine
    if device.type != "cuda":
        raise RuntimeError("radius_graph (B200 path) runs on CUDA only: pass CUDA positions or engine= "
                           "(there is no CPU fallback)")
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, float(cutoff))
    eng = _GRAPH_ENGINES.get(key)
    if eng is None:
        cfg = ModelConfig(32, 1, 4, float(cutoff), 1)
        state = {k: np.zeros(shape, dtype=np.float32) for k, shape in ckpt.expected_keys(cfg).items()}
        state.pop("rbf.centers"), state.pop("rbf.widths")
        state = ckpt.complete_state(state, cfg)     # reference formula for the RBF buffers
        eng = _GRAPH_ENGINES[key] = Engine(state, cfg, torch.device("cuda", index), "fp32")
    return eng


def radius_graph(positions: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None,
                 loop: bool = False, use_torch_cluster: bool = True, *, engine=None,
                 cell=None, pbc=None) -> torch.Tensor:
    """Drop-in for the reference ``radius_graph`` (student_model.py:165-191): ``[2,E]`` int64,
    row 0 = src, row 1 = dst, lexicographic order, computed by the CUDA neighbour kernels.
    Same call signature as the reference (``radius_graph(positions, r, batch)``): the CUDA context comes
    from a per-(device, cutoff) cache unless ``engine=`` supplies one; ``use_torch_cluster`` only selected
    the backend in the reference (same sorted edge set) and is accepted and ignored; ``loop=True`` adds
    the self pairs the reference keeps when it does not mask the diagonal (:101-103)."""
    if engine is None:
        engine = _graph_engine(positions.device, float(r))
    if abs(engine.cfg.cutoff - float(r)) > 0:
        raise ValueError("engine was created with a different cutoff")
    dev = engine.device
    pos = positions.detach().to(device=dev, dtype=torch.float32).contiguous()
    n = pos.shape[0]
    if batch is None:
        offsets, nb = torch.tensor([0, n], dtype=torch.int32, device=dev), 1
    else:
        offsets, nb = _offsets_from_batch(batch.to(dev))
    cells_d = pbc_d = None
    if cell is not None and pbc is not None and bool(torch.as_tensor(pbc).any()):
        check_minimum_image(cell, pbc, float(r), getattr(engine, "skin", 0.0))
        cells_d, pbc_d = StudentForceField.pack_cells(cell, pbc, nb, dev)
    eng = engine
    eng.ensure(n, nb)
    for _ in range(3):
        eng.neighbor_list_async(pos, offsets, nb, cells_d, pbc_d)
        st = eng.status()
        if not st.overflow:
            edges = eng.export_edges()
            if loop:   # self pairs, merged into the lexicographic (src, dst) order
                diag = torch.arange(n, dtype=torch.int64, device=dev)
                edges = torch.cat([edges, torch.stack([diag, diag])], dim=1)
                edges = edges[:, torch.argsort(edges[0] * n + edges[1])]
            return edges
        eng.reserve(n, int(st.num_edges * 1.25) + 64, nb)
    raise RuntimeError("edge workspace overflow persisted after growing")


__all__ = ["StudentForceField", "EnergyOnlyWrapper", "radius_graph", "check_minimum_image"]
