"""The oracle restatement against the golden vectors produced by the real reference
(tests/golden/make_golden.py) -- this is what pins the oracle (CPU only)."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden, load_weights
from oracle import painn_oracle as po
from oracle import reference_loader

KNOWN = {  # reference code + ONNX weights, FP32
    "original": {"h2o": -13.17164993, "benzene": -76.053284},
    "tiny": {"h2o": -15.46430206},
    "ultra_tiny": {"h2o": -17.01661682},
}
H2O_FORCES_ORIGINAL = np.array([[0.07642756, -0.25244409, 2.32016492],
                                [-0.02763237, 1.00698423, -1.25195813],
                                [-0.04879519, -0.75454021, -1.06820655]])


def test_known_answers(variant, golden):
    for case, e in KNOWN[variant].items():
        assert abs(float(golden[f"{case}_energy32"][0]) - e) < 2e-5
    if variant == "original":
        assert np.abs(golden["h2o_forces32"] - H2O_FORCES_ORIGINAL).max() < 2e-6


def test_oracle_matches_reference_outputs(variant, weights, golden):
    state, cfg = weights
    for case in golden_cases(golden):
        z, pos, off = golden[f"{case}_numbers"], golden[f"{case}_positions"], golden[f"{case}_offsets"]
        e, f = po.evaluate(state, cfg["cutoff"], z, pos, off)
        natoms = np.diff(off)
        assert np.max(np.abs(e - golden[f"{case}_energy32"]) / natoms) < 1e-6, case
        if case.endswith("_exact"):
            continue  # mirror-symmetric chain: forces ill-conditioned (see make_golden.py)
        assert np.max(np.abs(f - golden[f"{case}_forces32"])) < 2e-5, case


def test_oracle_fp64_matches_reference_fp64(variant, weights, golden):
    state, cfg = weights
    for case in ("h2o", "drug50", "isolated"):
        z, pos, off = golden[f"{case}_numbers"], golden[f"{case}_positions"], golden[f"{case}_offsets"]
        e, f = po.evaluate(state, cfg["cutoff"], z, pos, off, dtype=torch.float64)
        assert np.max(np.abs(e - golden[f"{case}_energy64"])) < 1e-9
        assert np.max(np.abs(f - golden[f"{case}_forces64"])) < 1e-9


def test_bruteforce_neighbor_list_equals_reference_graph(variant, weights, golden):
    _, cfg = weights
    for case in golden_cases(golden):
        ei, sh = po.neighbor_list(golden[f"{case}_positions"], golden[f"{case}_offsets"], cfg["cutoff"])
        assert np.array_equal(ei, golden[f"{case}_edge_index"]), case
        assert not sh.any()
        assert po.cutoff_ties(golden[f"{case}_positions"], ei, cfg["cutoff"]) == 0 or case.startswith("chain")


def test_periodic_neighbor_list_against_image_enumeration():
    """Minimum-image list == explicit enumeration of the 27 images (cubic cell, L >= 2 rc)."""
    rng = np.random.default_rng(5)
    L, rc, n = 11.0, 5.0, 60
    pos = rng.uniform(-3.0, L + 3.0, size=(n, 3)).astype(np.float32)  # deliberately unwrapped
    cell = np.eye(3) * L
    ei, sh = po.neighbor_list(pos, [0, n], rc, cell[None], np.array([[True, True, True]]))
    found = set()
    p64 = pos.astype(np.float64)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            d = p64[i] - p64[j]
            d -= L * np.round(d / L)
            if np.linalg.norm(d) <= rc - 1e-4:
                found.add((i, j))
    got = set(map(tuple, ei.T.tolist()))
    assert found <= got
    extra = got - found
    for i, j in extra:  # only pairs within 1e-4 of the cutoff may differ
        d = p64[i] - p64[j]
        d -= L * np.round(d / L)
        assert abs(np.linalg.norm(d) - rc) < 2e-4
    # symmetric, lexicographic, shift antisymmetric
    assert np.array_equal(ei, ei[:, np.lexsort((ei[1], ei[0]))])
    lookup = {(int(a), int(b)): k for k, (a, b) in enumerate(ei.T)}
    for k, (a, b) in enumerate(ei.T):
        assert np.array_equal(sh[lookup[(int(b), int(a))]], -sh[k])


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present")
def test_oracle_equals_live_reference_on_fresh_input(variant, weights):
    from mlff_distiller_b200 import synthetic
    from mlff_distiller_b200.checkpoint import infer_config
    state, cfg = weights
    model = reference_loader.build_reference_model(state, infer_config(state, cfg))
    structs = synthetic.druglike_batch(3, first=4242, ragged=True)
    z, pos, off = synthetic.concatenate(structs)
    pos = pos.astype(np.float32)
    batch = po.batch_from_offsets(off)
    p = torch.from_numpy(pos).requires_grad_(True)
    e_ref = model(torch.from_numpy(z), p, cell=None, pbc=None, batch=batch)
    f_ref = -torch.autograd.grad(e_ref, p, grad_outputs=torch.ones_like(e_ref))[0]
    e, f = po.evaluate(state, cfg["cutoff"], z, pos, off)
    assert np.max(np.abs(e - e_ref.detach().numpy())) < 1e-4
    assert np.max(np.abs(f - f_ref.numpy())) < 2e-5
