"""Strain derivative of the energy (virial / stress), SURVEY 8(f) rank 3.

The reference has no working stress (inference/ase_calculator.py:521-588 differentiates with respect
to a cell the model never reads and returns zeros), so parity is pinned two ways:
  * CPU: the oracle's autograd strain derivative against central finite differences of the oracle
    energy under a homogeneous deformation of positions and cell (FP64);
  * GPU: mlffd_virial against the oracle's autograd strain derivative on the same inputs.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_weights
from mlff_distiller_b200 import synthetic
from oracle import painn_oracle as po


def _periodic_case(n_mol=40, seed=11):
    box = synthetic.water_box(n_mol=n_mol, seed=seed, density=0.0334)   # L = 10.6 A >= 2 r_c
    # a sheared cell: the minimum image must be taken in fractional coordinates
    cell = box.cell.copy()
    cell[1, 0] = 0.7
    cell[2, 1] = -0.5
    return box.numbers, box.positions, cell, np.array([True, True, True])


def _oracle_virial(state, cfg, numbers, positions, offsets, cells=None, pbc=None, dtype=torch.float64):
    w = po.to_torch_weights(state, dtype)
    z = torch.from_numpy(np.asarray(numbers, dtype=np.int64))
    pos = torch.from_numpy(np.asarray(positions, dtype=np.float64)).to(dtype)
    batch = po.batch_from_offsets(offsets)
    ei, sh = po.neighbor_list(np.asarray(positions, dtype=np.float32), offsets, cfg["cutoff"], cells, pbc)
    shift_vec = None
    if cells is not None and pbc is not None and np.any(pbc):
        cells_t = torch.from_numpy(np.asarray(cells, dtype=np.float64)).to(dtype)
        shift_vec = torch.einsum("ek,ekc->ec", torch.from_numpy(sh).to(dtype), cells_t[batch[torch.from_numpy(ei[0])]])
    e, f, wv = po.energy_forces_virial(w, z, pos, cfg["cutoff"], batch, torch.from_numpy(ei), shift_vec)
    return np.atleast_1d(e.numpy()), f.numpy(), wv.numpy()


def test_oracle_virial_matches_finite_differences_periodic():
    state, cfg = load_weights("ultra_tiny")
    numbers, positions, cell, pbc = _periodic_case()
    offsets = [0, len(numbers)]
    _, _, w = _oracle_virial(state, cfg, numbers, positions, offsets, cell[None], pbc[None])
    h = 1e-5
    for a, c in ((0, 0), (1, 2), (2, 0)):
        es = []
        for sgn in (+1, -1):
            eps = np.zeros((3, 3))
            eps[a, c] = sgn * h
            defo = np.eye(3) + eps
            p2 = positions @ defo.T
            c2 = cell @ defo.T
            e, _ = po.evaluate(state, cfg["cutoff"], numbers, p2, offsets, c2[None], pbc[None], dtype=torch.float64,
                               dense_graph=False)
            es.append(float(e[0]))
        fd = (es[0] - es[1]) / (2 * h)
        assert abs(fd - w[0, a, c]) <= 1e-5 * max(1.0, abs(fd)), (a, c, fd, w[0, a, c])


def test_oracle_virial_open_boundary_equals_minus_sum_r_f():
    """Without periodic images the virial is -sum_i x_i (x) F_i (translation invariance)."""
    state, cfg = load_weights("tiny")
    s = synthetic.druglike_batch(1, first=77)[0]
    _, f, w = _oracle_virial(state, cfg, s.numbers, s.positions, [0, len(s.numbers)])
    ref = -np.einsum("ia,ic->ac", f, s.positions)
    assert np.max(np.abs(w[0] - ref)) <= 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["original", "tiny"])
def test_cuda_virial_matches_oracle(variant):
    from mlff_distiller_b200.student_model import StudentForceField
    state, cfg = load_weights(variant)
    model = StudentForceField.load(GOLDEN / f"weights_{variant}.npz", device="cuda:0", pbc_mode="minimum_image")
    dev = "cuda:0"
    # (1) ragged open-boundary batch
    structs = synthetic.druglike_batch(6, first=40, ragged=True) + [synthetic.water()]
    z, pos, off = synthetic.concatenate(structs)
    pos32 = pos.astype(np.float32)
    z_d = torch.from_numpy(z.astype(np.int32)).to(dev)
    p_d = torch.from_numpy(pos32).to(dev)
    o_d = torch.from_numpy(off.astype(np.int32)).to(dev)
    e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(structs))
    w = model.virial_of_last_call(o_d, len(structs)).cpu().numpy()
    _, f_ref, w_ref = _oracle_virial(state, cfg, z, pos32.astype(np.float64), off)
    scale = max(1.0, float(np.abs(w_ref).max()))
    assert np.max(np.abs(w - w_ref)) <= 2e-5 * scale, np.max(np.abs(w - w_ref))
    # translation invariance per structure, from the CUDA forces themselves
    for b in range(len(structs)):
        sl = slice(off[b], off[b + 1])
        ref = -np.einsum("ia,ic->ac", f.cpu().numpy()[sl].astype(np.float64), pos32[sl].astype(np.float64))
        assert np.max(np.abs(w[b] - ref)) <= 5e-4 * max(1.0, np.abs(ref).max())
    # (2) periodic, sheared cell, minimum image
    numbers, positions, cell, pbc = _periodic_case()
    zt = torch.from_numpy(numbers)
    pt = torch.from_numpy(positions.astype(np.float32))
    e1, f1, stress = model.predict_energy_forces_stress(zt, pt, torch.from_numpy(cell), torch.from_numpy(pbc))
    _, _, w_ref = _oracle_virial(state, cfg, numbers, positions.astype(np.float32).astype(np.float64),
                                 [0, len(numbers)], cell[None], pbc[None])
    vol = abs(np.linalg.det(cell))
    s_ref = 0.5 * (w_ref[0] + w_ref[0].T) / vol
    assert np.max(np.abs(stress.cpu().numpy() - s_ref)) <= 2e-5 * max(1.0, np.abs(s_ref).max())
    # energy-only call invalidates the edge adjoints: asking for the virial must fail loudly
    model.forward(zt, pt, torch.from_numpy(cell), torch.from_numpy(pbc))
    with pytest.raises(RuntimeError):
        model.virial_of_last_call(torch.tensor([0, len(numbers)], dtype=torch.int32, device=dev), 1)


@pytest.mark.gpu
def test_calculator_stress_voigt():
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    state, cfg = load_weights("ultra_tiny")
    calc = StudentForceFieldCalculator(GOLDEN / "weights_ultra_tiny.npz", device="cuda", pbc_mode="minimum_image",
                                       enable_stress=True)
    numbers, positions, cell, pbc = _periodic_case(n_mol=48, seed=3)
    box = synthetic.Structure(numbers, positions, cell=cell, pbc=pbc)
    calc.calculate(box, properties=["energy", "forces", "stress"])
    s = calc.results["stress"]
    assert s.shape == (6,)
    _, _, w_ref = _oracle_virial(state, cfg, numbers, positions.astype(np.float32).astype(np.float64),
                                 [0, len(numbers)], cell[None], pbc[None])
    sig = 0.5 * (w_ref[0] + w_ref[0].T) / abs(np.linalg.det(cell))
    voigt = np.array([sig[0, 0], sig[1, 1], sig[2, 2], sig[1, 2], sig[0, 2], sig[0, 1]])
    assert np.max(np.abs(s - voigt)) <= 2e-5 * max(1.0, np.abs(voigt).max())
    # the single-structure branch of calculate_batch hands the stress out like the reference's (:626-633)
    one = calc.calculate_batch([box], properties=["energy", "forces", "stress"])
    assert np.max(np.abs(one[0]["stress"] - voigt)) <= 2e-5 * max(1.0, np.abs(voigt).max())
    # no cell -> zeros, like the reference (:548-550); a cell without a periodic axis -> zeros as well
    mol = synthetic.water()
    calc.calculate(mol, properties=["energy", "forces", "stress"])
    assert np.array_equal(calc.results["stress"], np.zeros(6))
    open_box = synthetic.Structure(numbers, positions, cell=cell, pbc=np.zeros(3, dtype=bool))
    calc.calculate(open_box, properties=["energy", "forces", "stress"])
    assert np.array_equal(calc.results["stress"], np.zeros(6))
    # pbc_mode='ignore' reproduces the reference, whose model never reads the cell: its stress is zeros
    ref_mode = StudentForceFieldCalculator(GOLDEN / "weights_ultra_tiny.npz", device="cuda", enable_stress=True)
    ref_mode.calculate(box, properties=["energy", "forces", "stress"])
    assert np.array_equal(ref_mode.results["stress"], np.zeros(6))
