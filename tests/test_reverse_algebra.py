"""The identities behind the three kinds of edge-adjoint slab (csrc/readout.cuh, csrc/message_spline.cuh,
DESIGN section 2 items 4-6), checked in FP64 on the CPU against the oracle's autograd -- no GPU involved.

With (u_bar_e, d_bar_e) the adjoint of edge e = (j -> i) (unit vector and distance-through-the-filter),
r_bar_e = P(g_e, u_bar_e, d_bar_e) its position adjoint (``edge_position_adjoint``) and rev(e) = (i -> j):
  * forces      F_k = sum_{e -> k} (r_bar_e - r_bar_rev(e))
  * pair slab   r_bar_e - r_bar_rev(e) = P(g_e, u_bar_e - u_bar_rev(e), d_bar_e + d_bar_rev(e)), stored at the
                upper entry of the pair only
  * swapped slab: what is stored at e belongs to rev(e)
  * virial      W = sum_e r_bar_e (x) r_e, with the same three readings.
The test splits the oracle's total adjoints into three parts, stores them the three ways, runs a line-by-line
numpy restatement of force_kernel / virial_kernel over the mixed slabs and compares with -dE/dx and dE/d(strain)
from autograd."""
import numpy as np
import torch

import oracle.painn_oracle as po
from conftest import load_weights
from mlff_distiller_b200 import synthetic

EPS = 1e-8   # kUnitEps: unit = r / (d + eps), student_model.py:715


def position_adjoint(g, adj):
    """readout.cuh:edge_position_adjoint for arrays: g = (u, d) [E,4], adj = (u_bar, d_bar) [E,4]."""
    u, d = g[:, :3], g[:, 3]
    q = d + EPS
    udot = np.einsum("ea,ea->e", u, adj[:, :3])
    scale = np.where(d > 0, (adj[:, 3] - udot / q) * (q / np.where(d > 0, d, 1.0)), 0.0)
    return scale[:, None] * u + adj[:, :3] / q[:, None]


def test_three_slab_kinds_reproduce_autograd_forces_and_virial():
    state, cfg = load_weights("ultra_tiny")
    structs = synthetic.druglike_batch(3, first=77, ragged=True)
    z, pos, off = synthetic.concatenate(structs)
    w = po.to_torch_weights(state, torch.float64)
    batch = po.batch_from_offsets(off)
    zt, pt = torch.from_numpy(np.asarray(z, np.int64)), torch.from_numpy(pos.astype(np.float32)).double()
    e, f, keep = po.energy_and_forces_with_adjoints(w, zt, pt, cfg["cutoff"], batch)
    _, _, w_ref = po.energy_forces_virial(w, zt, pt, cfg["cutoff"], batch)
    src, dst = (t.numpy() for t in keep["edge_index"])
    # CSR order of the CUDA path: row = destination, sources ascending
    order = np.lexsort((src, dst))
    src, dst = src[order], dst[order]
    unit = keep["unit"].detach().numpy()[order]
    dist = keep["d"].detach().numpy()[order]
    ubar = keep["unit"].grad.numpy()[order]
    # autograd's d.grad also contains the path through unit = r / (d + eps); the kernels carry d_bar through
    # the filter only and let the position adjoint handle the normalisation
    dbar = keep["d"].grad.numpy()[order] + np.einsum("ea,ea->e", ubar, unit) / (dist + EPS)
    E = len(src)
    key = {(int(s), int(t)): k for k, (s, t) in enumerate(zip(src, dst))}
    rev = np.array([key[(int(t), int(s))] for s, t in zip(src, dst)])
    assert np.array_equal(rev[rev], np.arange(E)) and np.allclose(unit[rev], -unit) and np.array_equal(dist[rev], dist)
    geo = np.concatenate([unit, dist[:, None]], axis=1)
    adj = np.concatenate([ubar, dbar[:, None]], axis=1)

    # --- plain statement: F_k = sum_{e -> k} (r_bar_e - r_bar_rev(e)) ---
    rbar = position_adjoint(geo, adj)
    forces = np.zeros((len(z), 3))
    np.add.at(forces, dst, rbar - rbar[rev])
    assert np.abs(forces - f.numpy()).max() <= 1e-9 * max(1.0, np.abs(f.numpy()).max())

    # --- three slab kinds: 30 % as pair sums, 25 % direct, 45 % swapped ---
    upper = rev > np.arange(E)
    a_pair, a_direct, a_swapped = 0.30 * adj, 0.25 * adj, 0.45 * adj
    pair_slab = np.full((E, 4), np.nan)                      # lower entries are never written
    pair_slab[upper, :3] = a_pair[upper, :3] - a_pair[rev[upper], :3]
    pair_slab[upper, 3] = a_pair[upper, 3] + a_pair[rev[upper], 3]
    direct_slab = a_direct
    swapped_slab = a_swapped[rev]                            # entry e holds the adjoint of rev(e)

    # force_kernel
    adj_e, adj_r = direct_slab.copy(), direct_slab[rev].copy()
    adj_e[upper] += pair_slab[upper]
    lower = ~upper
    adj_r[lower] += pair_slab[rev[lower]]
    adj_e += swapped_slab[rev]
    adj_r += swapped_slab
    a, b = position_adjoint(geo, adj_e), position_adjoint(geo[rev], adj_r)
    forces3 = np.zeros((len(z), 3))
    np.add.at(forces3, dst, a - b)
    assert np.abs(forces3 - f.numpy()).max() <= 1e-9 * max(1.0, np.abs(f.numpy()).max())

    # virial_kernel
    r_vec = unit * (dist + EPS)[:, None]
    acc = direct_slab.copy()
    acc[upper] += pair_slab[upper]
    rb = position_adjoint(geo, acc)
    geo_rev = np.concatenate([-unit, dist[:, None]], axis=1)
    rb -= position_adjoint(geo_rev, swapped_slab)
    per_edge = np.einsum("ea,eb->eab", rb, r_vec)
    struct_of_edge = batch.numpy()[dst]
    virial = np.zeros((len(off) - 1, 3, 3))
    np.add.at(virial, struct_of_edge, per_edge)
    assert np.abs(virial - w_ref.numpy()).max() <= 1e-9 * max(1.0, np.abs(w_ref.numpy()).max())
