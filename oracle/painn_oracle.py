"""ORACLE (test infrastructure, not product code): CPU restatement of the PaiNN-student E+F path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product path (mlff_distiller_b200) never does.

Parity status: PINNED for open boundaries -- ``tests/golden/make_golden.py`` ran the imported
reference (``/root/reference/src/mlff_distiller/models/student_model.py``, loaded by file path)
on the committed inputs and this restatement reproduces those energies/forces bit-for-bit on the
same torch build (tests/test_oracle_golden.py).  Periodic boundaries: parity UNPINNED -- the
reference silently ignores ``cell``/``pbc`` (student_model.py:694-703), so the minimum-image
semantics in :func:`neighbor_list` are this project's definition.

Every function cites the reference lines it restates (paths relative to /root/reference).
The arithmetic deliberately uses the same ATen ops in the same order as the reference so the
restatement and the reference agree to the last bit on CPU.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# neighbour list
# --------------------------------------------------------------------------------------

def radius_graph_dense(positions: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None
                       ) -> torch.Tensor:
    """Restates ``radius_graph_native`` (student_model.py:63-109): dense [N,N] mask, inclusive
    ``<=``, same-structure pairs, no self loops, edges in lexicographic (src, dst) order."""
    n = positions.shape[0]
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    diff = positions.unsqueeze(0) - positions.unsqueeze(1)
    dist = torch.norm(diff, dim=2)
    mask = (batch.unsqueeze(0) == batch.unsqueeze(1)) & (dist <= r)
    mask = mask & ~torch.eye(n, dtype=torch.bool)
    src, dst = torch.where(mask)
    return torch.stack([src, dst], dim=0)


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def neighbor_list(positions: np.ndarray, offsets: Sequence[int], cutoff: float,
                  cells: Optional[np.ndarray] = None, pbc: Optional[np.ndarray] = None,
                  chunk: int = 1024) -> Tuple[np.ndarray, np.ndarray]:
    """Brute-force neighbour list with the FIXED FP32 operation order the CUDA kernels use.

    Open boundaries follow the semantics of ``radius_graph_native`` (student_model.py:89-107):
    ordered pairs i != j of the same structure with ``d <= cutoff`` (inclusive).  The distance is
    evaluated as ``sqrt((dx*dx + dy*dy) + dz*dz)`` with every product and sum rounded to FP32
    (no FMA), dx = x_src - x_dst.  ``torch.norm`` is not reproducible at the ulp level, so pairs
    with ``|d - cutoff| <= 4 ulp`` are ties (tests count them).

    Periodic boundaries (project-defined, parity unpinned): per pair, with D = x_src - x_dst,
    ``f_k = (D_x*inv[0,k] + D_y*inv[1,k]) + D_z*inv[2,k]``, ``n_k = rint(f_k)`` on periodic axes
    (0 elsewhere), ``D' = D - ((n_0*c[0] + n_1*c[1]) + n_2*c[2])``; all FP32, no FMA; ``inv`` is the
    FP64 inverse of the cell rounded to FP32.  No self-image edges.

    Returns ``(edge_index [2,E] int64 lexicographic (src,dst), shifts [E,3] int32)`` where the
    edge vector is ``x_src - x_dst - shifts @ cell``.
    """
    pos = _f32(positions)
    offsets = np.asarray(offsets, dtype=np.int64)
    srcs: List[np.ndarray] = []
    dsts: List[np.ndarray] = []
    shifts: List[np.ndarray] = []
    rc = np.float32(cutoff)
    for b in range(len(offsets) - 1):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        p = pos[lo:hi]
        periodic = pbc is not None and bool(np.any(pbc[b]))
        if periodic:
            cell = _f32(cells[b])
            inv = _f32(np.linalg.inv(np.asarray(cells[b], dtype=np.float64)))
            per = np.asarray(pbc[b], dtype=bool)
        for s0 in range(0, hi - lo, chunk):
            s1 = min(hi - lo, s0 + chunk)
            d = p[s0:s1, None, :] - p[None, :, :]  # [src chunk, dst, 3] = x_src - x_dst
            dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
            if periodic:
                n = np.zeros(d.shape, dtype=np.float32)
                for k in range(3):
                    if per[k]:
                        f = (dx * inv[0, k] + dy * inv[1, k]) + dz * inv[2, k]
                        n[..., k] = np.rint(f)
                sx = (n[..., 0] * cell[0, 0] + n[..., 1] * cell[1, 0]) + n[..., 2] * cell[2, 0]
                sy = (n[..., 0] * cell[0, 1] + n[..., 1] * cell[1, 1]) + n[..., 2] * cell[2, 1]
                sz = (n[..., 0] * cell[0, 2] + n[..., 1] * cell[1, 2]) + n[..., 2] * cell[2, 2]
                dx, dy, dz = dx - sx, dy - sy, dz - sz
            dist = np.sqrt((dx * dx + dy * dy) + dz * dz)
            mask = dist <= rc
            ii = np.arange(s0, s1)
            mask[ii - s0, ii] = False
            si, di = np.nonzero(mask)
            srcs.append(si + s0 + lo)
            dsts.append(di + lo)
            if periodic:
                shifts.append(n[si, di].astype(np.int32))
            else:
                shifts.append(np.zeros((len(si), 3), dtype=np.int32))
    if srcs:
        ei = np.stack([np.concatenate(srcs), np.concatenate(dsts)]).astype(np.int64)
        sh = np.concatenate(shifts).astype(np.int32)
    else:
        ei = np.zeros((2, 0), dtype=np.int64)
        sh = np.zeros((0, 3), dtype=np.int32)
    return ei, sh


def cutoff_ties(positions: np.ndarray, edge_index: np.ndarray, cutoff: float, ulps: int = 4) -> int:
    """Number of listed pairs whose FP32 distance lies within ``ulps`` of the cutoff."""
    pos = _f32(positions)
    d = pos[edge_index[0]] - pos[edge_index[1]]
    dist = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    return int(np.sum(np.abs(dist - np.float32(cutoff)) <= ulps * np.spacing(np.float32(cutoff))))


# --------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------

def to_torch_weights(state: Mapping[str, np.ndarray], dtype=torch.float32) -> Dict[str, torch.Tensor]:
    return {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in state.items()}


def num_layers(w: Mapping[str, torch.Tensor]) -> int:
    return len({int(k.split(".")[1]) for k in w if k.startswith("interactions.")})


def edge_features(w, edge_vector: torch.Tensor, cutoff: float):
    """student_model.py:711-722 with GaussianRBF.forward (:249-255) and CosineCutoff.forward
    (:285-292) inlined."""
    d = torch.norm(edge_vector, dim=1)
    unit = edge_vector / (d.unsqueeze(1) + 1e-8)
    diff = d.unsqueeze(-1) - w["rbf.centers"]
    gamma = 1.0 / (w["rbf.widths"] ** 2)
    rbf = torch.exp(-gamma * diff ** 2)
    fc = 0.5 * (torch.cos(np.pi * d / cutoff) + 1.0)
    fc = fc * (d < cutoff).to(d.dtype)
    return d, unit, rbf * fc.unsqueeze(-1)


def message(w, l: int, s, v, edge_index, edge_rbf, unit):
    """PaiNNMessage.forward (student_model.py:346-387)."""
    p = f"interactions.{l}.message.rbf_to_scalar."
    src, dst = edge_index
    h = s.shape[1]
    filt = F.linear(F.silu(F.linear(edge_rbf, w[p + "0.weight"], w[p + "0.bias"])),
                    w[p + "2.weight"], w[p + "2.bias"])
    fa, fb, fc = torch.split(filt, h, dim=-1)
    s_out = torch.zeros_like(s)
    s_out.index_add_(0, dst, s[src] * fa)
    s_out = s + s_out
    vm = v[src] * fb.unsqueeze(1) + unit.unsqueeze(-1) * fc.unsqueeze(1)
    v_out = torch.zeros_like(v)
    v_out.index_add_(0, dst, vm)
    v_out = v + v_out
    return s_out, v_out, filt


def update(w, l: int, s, v):
    """PaiNNUpdate.forward (student_model.py:434-470); ``mixing_matrix`` acts on the xyz axis."""
    p = f"interactions.{l}.update."
    h = s.shape[1]
    norms = torch.norm(v, dim=1)
    y1 = F.linear(torch.cat([s, norms], dim=-1), w[p + "update_mlp.0.weight"],
                  w[p + "update_mlp.0.bias"])
    upd = F.linear(F.silu(y1), w[p + "update_mlp.2.weight"], w[p + "update_mlp.2.bias"])
    ds, g1, g2 = torch.split(upd, h, dim=-1)
    mix = w.get(p + "mixing_matrix")
    if mix is None:  # pruned by the ONNX export for the last layer: dead for E and F (:736)
        mix = torch.zeros(3, 3, dtype=v.dtype)
    mixed = torch.einsum("ij,njk->nik", mix, v)
    return s + ds, v * g1.unsqueeze(1) + mixed * g2.unsqueeze(1), y1, g1, g2


def readout(w, s):
    """energy_head (student_model.py:599-605, 736)."""
    x = F.silu(F.linear(s, w["energy_head.0.weight"], w["energy_head.0.bias"]))
    x = F.silu(F.linear(x, w["energy_head.2.weight"], w["energy_head.2.bias"]))
    return F.linear(x, w["energy_head.4.weight"], w["energy_head.4.bias"])


def forward(w, atomic_numbers: torch.Tensor, positions: torch.Tensor, cutoff: float,
            batch: Optional[torch.Tensor] = None, edge_index: Optional[torch.Tensor] = None,
            shift_vectors: Optional[torch.Tensor] = None, keep: Optional[dict] = None,
            strain: Optional[torch.Tensor] = None):
    """StudentForceField.forward (student_model.py:675-757).

    ``edge_index`` defaults to the dense open-boundary graph; a periodic caller injects its own
    edge list and Cartesian ``shift_vectors`` (edge vector = x_src - x_dst - shift).
    ``keep`` (a dict) receives every intermediate, with ``retain_grad`` when differentiable, so
    tests can compare the CUDA stages and adjoints one by one.
    ``strain`` ``[B,3,3]`` deforms every edge vector of structure b as r -> (1 + eps_b) r (positions
    and cell strained together), the handle :func:`energy_forces_virial` differentiates through.
    """
    n = atomic_numbers.shape[0]
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    s = w["embedding.weight"][atomic_numbers]
    hdim = s.shape[1]
    v = torch.zeros(n, 3, hdim, dtype=positions.dtype)
    if edge_index is None:
        edge_index = radius_graph_dense(positions, cutoff, batch)
    src, dst = edge_index
    edge_vector = positions[src] - positions[dst]
    if shift_vectors is not None:
        edge_vector = edge_vector - shift_vectors
    if strain is not None:
        edge_vector = edge_vector + torch.einsum("eab,eb->ea", strain[batch[src]], edge_vector)
    d, unit, edge_rbf = edge_features(w, edge_vector, cutoff)

    def _keep(name, t):
        if keep is not None:
            if t.requires_grad:
                t.retain_grad()
            keep[name] = t
        return t

    _keep("edge_index", edge_index)
    _keep("d", d)
    _keep("unit", unit)
    _keep("edge_rbf", edge_rbf)
    for l in range(num_layers(w)):
        _keep(f"s_in{l}", s)
        _keep(f"v_in{l}", v)
        s, v, filt = message(w, l, s, v, edge_index, edge_rbf, unit)
        _keep(f"filter{l}", filt)
        _keep(f"s_msg{l}", s)
        _keep(f"v_msg{l}", v)
        s, v, y1, g1, g2 = update(w, l, s, v)
        _keep(f"y1_{l}", y1)
        _keep(f"g1_{l}", g1)
        _keep(f"g2_{l}", g2)
    _keep("s_out", s)
    eps = _keep("atomic_energies", readout(w, s))
    if batch.numel() == 0 or batch.max() == 0:
        return torch.sum(eps)
    nb = int(batch.max()) + 1
    total = torch.zeros(nb, dtype=eps.dtype)
    for i in range(nb):
        total[i] = eps[batch == i].sum()
    return total


def energy_and_forces(w, atomic_numbers, positions, cutoff: float, batch=None, edge_index=None,
                      shift_vectors=None, keep: Optional[dict] = None):
    """predict_energy_and_forces (student_model.py:782-795) / _batch_forward
    (inference/ase_calculator.py:757-763): F = -dE/dx by reverse-mode autograd."""
    pos = positions.detach().clone().requires_grad_(True)
    e = forward(w, atomic_numbers, pos, cutoff, batch, edge_index, shift_vectors, keep)
    grad = torch.autograd.grad(e, pos, grad_outputs=torch.ones_like(e))[0]
    return e.detach(), -grad


def energy_forces_virial(w, atomic_numbers, positions, cutoff: float, batch=None, edge_index=None,
                         shift_vectors=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(E [B] or scalar, F [N,3], W [B,3,3]) with W_ac = dE_b / d eps_ac at eps = 0 by autograd.
    stress = sym(W) / V.  The reference has no working counterpart: its _compute_stress
    (inference/ase_calculator.py:521-588) asks autograd for dE/d(cell) of a model that never reads
    the cell, catches the error and returns zeros -- so this function is the definition the CUDA
    virial is tested against, together with finite differences of the energy (tests)."""
    n = atomic_numbers.shape[0]
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    nb = int(batch.max()) + 1 if n else 1
    pos = positions.detach().clone().requires_grad_(True)
    eps = torch.zeros(nb, 3, 3, dtype=positions.dtype, requires_grad=True)
    e = forward(w, atomic_numbers, pos, cutoff, batch, edge_index, shift_vectors, None, eps)
    gpos, geps = torch.autograd.grad(e.sum(), (pos, eps))
    return e.detach(), -gpos, geps


def energy_and_forces_with_adjoints(w, atomic_numbers, positions, cutoff: float, batch=None,
                                    edge_index=None, shift_vectors=None) -> Tuple[torch.Tensor, torch.Tensor, dict]:
    """Same as :func:`energy_and_forces` but returns every intermediate and its adjoint
    (``keep[name].grad``) via ``backward``; used by the stage-level parity tests."""
    keep: dict = {}
    pos = positions.detach().clone().requires_grad_(True)
    e = forward(w, atomic_numbers, pos, cutoff, batch, edge_index, shift_vectors, keep)
    e.sum().backward()
    return e.detach(), -pos.grad.detach(), keep


def batch_from_offsets(offsets: Sequence[int]) -> torch.Tensor:
    offsets = np.asarray(offsets, dtype=np.int64)
    return torch.from_numpy(np.repeat(np.arange(len(offsets) - 1), np.diff(offsets)))


def evaluate(state: Mapping[str, np.ndarray], cutoff: float, numbers: np.ndarray,
             positions: np.ndarray, offsets: Optional[Sequence[int]] = None,
             cells: Optional[np.ndarray] = None, pbc: Optional[np.ndarray] = None,
             dtype=torch.float32, dense_graph: bool = True):
    """Convenience wrapper: numpy in, numpy out.

    Open boundaries with ``dense_graph=True`` go through :func:`radius_graph_dense` (the
    reference's own graph builder); otherwise the brute-force FP32 list of :func:`neighbor_list`
    is injected (required for periodic systems and for structures too large for the dense mask).
    Returns ``(energies [B] float64, forces [N,3] float64)``.
    """
    numbers = np.asarray(numbers)
    n = len(numbers)
    if offsets is None:
        offsets = [0, n]
    w = to_torch_weights(state, dtype)
    z = torch.from_numpy(numbers.astype(np.int64))
    pos32 = _f32(positions)
    pos = torch.from_numpy(np.asarray(positions if dtype == torch.float64 else pos32)).to(dtype)
    batch = batch_from_offsets(offsets)
    periodic = pbc is not None and bool(np.any(pbc))
    if periodic or not dense_graph:
        ei, sh = neighbor_list(pos32, offsets, cutoff, cells, pbc)
        shift_vec = None
        if periodic:
            cells_t = torch.from_numpy(np.asarray(cells, dtype=np.float64)).to(dtype)
            eb = batch[torch.from_numpy(ei[0])]
            shift_vec = torch.einsum("ek,ekc->ec", torch.from_numpy(sh).to(dtype), cells_t[eb])
        e, f = energy_and_forces(w, z, pos, cutoff, batch, torch.from_numpy(ei), shift_vec)
    else:
        e, f = energy_and_forces(w, z, pos, cutoff, batch)
    return np.atleast_1d(e.double().numpy()), f.double().numpy()
