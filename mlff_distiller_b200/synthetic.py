"""Seeded synthetic structures for the BASELINE.json configs.

There is no network for datasets; every workload is generated from fixed seeds so the CUDA path,
the oracle and the CPU baseline see identical inputs.  Generators assert a minimum pair
separation of 0.95 Angstrom (physically sane inputs keep FP32 noise below the force tolerance).

* C1 ``water()``: ASE g2 H2O geometry.
* C2 ``druglike_batch()``: ~50-atom random organic-like blobs at 0.08 atoms/A^3.
* C3 ``alkane_chain()``: the reference's "peptide" generator
  (scripts/peptide_folding_simulation.py:50-71: carbon chain, 1.5 A spacing, two H at +-1.1 A)
  extended to 100 CH2 units = 300 atoms.
* C4 ``water_box()``: 3333 rigid waters in a cubic periodic box at 0.0334 molecules/A^3.
* C5 ``druglike_batch(ragged=True)``: ragged sizes 20..80 for the screening sweep.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

ELEMENTS = np.array([1, 6, 7, 8, 9, 16], dtype=np.int64)
ELEMENT_P = np.array([0.45, 0.35, 0.08, 0.10, 0.01, 0.01])
MIN_SEPARATION = 0.95


class Structure:
    """Minimal Atoms-like record (duck-types the ASE getters the calculator reads)."""

    def __init__(self, numbers, positions, cell=None, pbc=None, masses=None):
        self.numbers = np.asarray(numbers, dtype=np.int64)
        self.positions = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
        self.cell = np.zeros((3, 3)) if cell is None else np.asarray(cell, dtype=np.float64)
        self.pbc = np.zeros(3, dtype=bool) if pbc is None else np.asarray(pbc, dtype=bool)
        self.calc = None
        self._masses = masses

    def __len__(self):
        return len(self.numbers)

    def get_positions(self):
        return self.positions.copy()

    def set_positions(self, p):
        self.positions = np.asarray(p, dtype=np.float64).reshape(-1, 3)

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_cell(self):
        return self.cell.copy()

    def get_pbc(self):
        return self.pbc.copy()

    def get_masses(self):
        if self._masses is not None:
            return np.asarray(self._masses, dtype=np.float64)
        from .md import ATOMIC_MASSES
        return ATOMIC_MASSES[self.numbers]

    def copy(self):
        return Structure(self.numbers, self.positions.copy(), self.cell.copy(), self.pbc.copy(),
                         self._masses)

    def get_potential_energy(self):
        return self.calc.get_potential_energy(self)

    def get_forces(self):
        return self.calc.get_forces(self)

    def get_stress(self):
        return self.calc.get_stress(self)


def water() -> Structure:
    """C1: ASE g2 H2O (O at z=0.119262, H at y=+-0.763239, z=-0.477047)."""
    pos = np.array([[0.0, 0.0, 0.119262], [0.0, 0.763239, -0.477047], [0.0, -0.763239, -0.477047]])
    return Structure([8, 1, 1], pos)


def benzene() -> Structure:
    """ASE g2 C6H6 geometry (planar, C-C 1.395 A ring radius, C-H 1.087 A)."""
    pos = np.array([
        [0.0, 1.395248, 0.0], [1.20832, 0.697624, 0.0], [1.20832, -0.697624, 0.0],
        [0.0, -1.395248, 0.0], [-1.20832, -0.697624, 0.0], [-1.20832, 0.697624, 0.0],
        [0.0, 2.482360, 0.0], [2.149787, 1.241180, 0.0], [2.149787, -1.241180, 0.0],
        [0.0, -2.482360, 0.0], [-2.149787, -1.241180, 0.0], [-2.149787, 1.241180, 0.0]])
    return Structure([6] * 6 + [1] * 6, pos)


def _blob_positions(rng: np.random.Generator, n: int, density: float = 0.08,
                    min_sep: float = MIN_SEPARATION) -> np.ndarray:
    """Rejection-sample n points uniformly in a sphere of the given number density."""
    radius = (3.0 * n / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    pts = np.empty((n, 3))
    k = 0
    min2 = min_sep * min_sep
    while k < n:
        cand = rng.uniform(-radius, radius, size=3)
        if cand @ cand > radius * radius:
            continue
        if k and np.min(np.sum((pts[:k] - cand) ** 2, axis=1)) < min2:
            continue
        pts[k] = cand
        k += 1
    return pts


def druglike(seed: int, n: int = 50) -> Structure:
    """One C2/C5 structure: ``default_rng(seed)``; elements from {H,C,N,O,F,S}."""
    rng = np.random.default_rng(seed)
    z = rng.choice(ELEMENTS, size=n, p=ELEMENT_P)
    pos = _blob_positions(rng, n)
    return Structure(z, pos)


def druglike_batch(count: int, first: int = 0, n: int = 50, ragged: bool = False
                   ) -> List[Structure]:
    """C2 (fixed n) or C5 (``ragged``: n ~ U{20..80}) batches; structure s uses seed 1000+s."""
    out = []
    for s in range(first, first + count):
        if ragged:
            ns = int(np.random.default_rng(500000 + s).integers(20, 81))
        else:
            ns = n
        out.append(druglike(1000 + s, ns))
    return out


def alkane_chain(units: int = 100) -> Structure:
    """C3: straight CH2 chain, 3*units atoms, along x with 1.5 A spacing."""
    pos, z = [], []
    for i in range(units):
        x = 1.5 * i
        pos += [[x, 0.0, 0.0], [x, 1.1, 0.0], [x, -1.1, 0.0]]
        z += [6, 1, 1]
    return Structure(z, np.array(pos))


def water_box(n_mol: int = 3333, seed: int = 3000, density: float = 0.0334,
              jitter: float = 0.3) -> Structure:
    """C4: rigid waters (0.9572 A, 104.52 deg) at random orientation on a jittered cubic grid,
    cubic periodic box L = (n_mol / density)^(1/3) (46.383 A for 3333 molecules)."""
    rng = np.random.default_rng(seed)
    L = (n_mol / density) ** (1.0 / 3.0)
    g = int(np.ceil(n_mol ** (1.0 / 3.0)))
    idx = rng.permutation(g ** 3)[:n_mol]
    grid = np.stack(np.unravel_index(idx, (g, g, g)), axis=1).astype(np.float64)
    o = (grid + 0.5) * (L / g) + rng.normal(0.0, jitter, size=(n_mol, 3))
    ang = np.deg2rad(104.52)
    h1 = 0.9572 * np.array([np.sin(ang / 2), 0.0, np.cos(ang / 2)])
    h2 = 0.9572 * np.array([-np.sin(ang / 2), 0.0, np.cos(ang / 2)])
    # random rotations from normalised quaternions
    q = rng.normal(size=(n_mol, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, zq = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([
        np.stack([1 - 2 * (y * y + zq * zq), 2 * (x * y - zq * w), 2 * (x * zq + y * w)], 1),
        np.stack([2 * (x * y + zq * w), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - x * w)], 1),
        np.stack([2 * (x * zq - y * w), 2 * (y * zq + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)
    pos = np.empty((n_mol, 3, 3))
    pos[:, 0] = o
    pos[:, 1] = o + R @ h1
    pos[:, 2] = o + R @ h2
    z = np.tile(np.array([8, 1, 1]), n_mol)
    return Structure(z, pos.reshape(-1, 3), cell=np.eye(3) * L, pbc=[True, True, True])


def concatenate(structures: Sequence[Structure]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(numbers int64 [N], positions float64 [N,3], offsets int64 [B+1]) of a structure list."""
    counts = np.array([len(s) for s in structures], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    z = np.concatenate([s.numbers for s in structures]) if structures else np.zeros(0, np.int64)
    pos = (np.concatenate([s.positions for s in structures]) if structures
           else np.zeros((0, 3)))
    return z, pos, offsets


def min_pair_distance(pos: np.ndarray, offsets: Optional[Sequence[int]] = None) -> float:
    pos = np.asarray(pos, dtype=np.float64)
    if offsets is None:
        offsets = [0, len(pos)]
    best = np.inf
    for b in range(len(offsets) - 1):
        p = pos[offsets[b]:offsets[b + 1]]
        if len(p) < 2:
            continue
        for s0 in range(0, len(p), 2048):
            d = np.linalg.norm(p[s0:s0 + 2048, None] - p[None], axis=-1)
            d[np.arange(len(d)), np.arange(s0, s0 + len(d))] = np.inf
            best = min(best, float(d.min()))
    return best
