"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances (BASELINE.json north_star, FP32): |dE| <= 1e-5 eV/atom, max|dF| <= 1e-4 eV/A; the
neighbour edge set must be bit-exact once sorted (it is produced already sorted).
A diagnostics JSON is written to gpurun_out/parity_<variant>.json on every run.
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_cases, load_golden, load_weights
from oracle import painn_oracle as po

pytestmark = pytest.mark.gpu

E_TOL = 1e-5   # eV / atom
F_TOL = 1e-4   # eV / A
OUT = ROOT / "gpurun_out"


def _model(variant, **kw):
    from mlff_distiller_b200.checkpoint import infer_config
    from mlff_distiller_b200.student_model import StudentForceField
    state, cfg = load_weights(variant)
    m = StudentForceField.from_state(state, infer_config(state, cfg), "cuda:0", **kw)
    return m, state, cfg


@pytest.fixture(scope="module")
def model_bundle(variant):
    return _model(variant)


def _run(model, z, pos, off, cells=None, pbc=None):
    dev = "cuda:0"
    z_d = torch.from_numpy(np.asarray(z, dtype=np.int32)).to(dev)
    p_d = torch.from_numpy(np.asarray(pos, dtype=np.float32)).to(dev)
    o_d = torch.from_numpy(np.asarray(off, dtype=np.int32)).to(dev)
    c_d = b_d = None
    if cells is not None:
        from mlff_distiller_b200.student_model import StudentForceField
        c_d, b_d = StudentForceField.pack_cells(torch.from_numpy(np.asarray(cells)),
                                                torch.from_numpy(np.asarray(pbc)), len(off) - 1, dev)
    e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(off) - 1, c_d, b_d)
    return e.double().cpu().numpy(), f.double().cpu().numpy()


def test_neighbor_list_bit_exact(model_bundle, golden, variant):
    model, state, cfg = model_bundle
    from mlff_distiller_b200.student_model import radius_graph
    eng = model.engine()
    for case in golden_cases(golden):
        pos, off = golden[f"{case}_positions"], golden[f"{case}_offsets"]
        batch = po.batch_from_offsets(off).cuda()
        ei = radius_graph(torch.from_numpy(pos), cfg["cutoff"], batch, engine=eng).cpu().numpy()
        assert ei.dtype == np.int64
        assert np.array_equal(ei, golden[f"{case}_edge_index"]), case


def test_neighbor_list_periodic_bit_exact(model_bundle):
    model, state, cfg = model_bundle
    from mlff_distiller_b200.student_model import radius_graph
    rng = np.random.default_rng(11)
    n, L = 400, 13.0
    pos = rng.uniform(-4.0, L + 4.0, size=(n, 3)).astype(np.float32)
    cell = np.array([[L, 0, 0], [1.5, L, 0], [0.7, -1.1, L + 2]], dtype=np.float64)  # triclinic
    for pbc in ([True, True, True], [True, False, True]):
        ei_ref, _ = po.neighbor_list(pos, [0, n], cfg["cutoff"], cell[None], np.array([pbc]))
        ei = radius_graph(torch.from_numpy(pos), cfg["cutoff"], None, engine=model.engine(),
                          cell=torch.from_numpy(cell), pbc=torch.tensor(pbc)).cpu().numpy()
        assert np.array_equal(ei, ei_ref), pbc


def test_filter_table_matches_oracle(model_bundle):
    model, state, cfg = model_bundle
    eng = model.engine()
    w64 = po.to_torch_weights(state, torch.float64)
    rc = cfg["cutoff"]
    d = torch.cat([torch.linspace(0.4, rc + 0.3, 2000), torch.tensor([rc, rc - 1e-6, 0.9572])]).float()
    for l in range(cfg["num_interactions"]):
        f, df = eng.filter_table(l, d)
        dd = d.double().requires_grad_(True)
        vec = torch.stack([dd, torch.zeros_like(dd), torch.zeros_like(dd)], dim=1)
        _, _, rbf = po.edge_features(w64, vec, rc)
        p = f"interactions.{l}.message.rbf_to_scalar."
        filt = torch.nn.functional.linear(
            torch.nn.functional.silu(torch.nn.functional.linear(rbf, w64[p + "0.weight"], w64[p + "0.bias"])),
            w64[p + "2.weight"], w64[p + "2.bias"])
        # derivative of every output channel w.r.t. its own distance via a forward-mode trick
        jac = torch.autograd.functional.jvp(
            lambda x: torch.nn.functional.linear(
                torch.nn.functional.silu(torch.nn.functional.linear(
                    po.edge_features(w64, torch.stack([x, torch.zeros_like(x), torch.zeros_like(x)], 1), rc)[2],
                    w64[p + "0.weight"], w64[p + "0.bias"])),
                w64[p + "2.weight"], w64[p + "2.bias"]),
            dd.detach(), torch.ones_like(dd))[1]
        scale = float(filt.abs().max())
        assert float((f.double().cpu() - filt.detach()).abs().max()) < 2e-5 * max(scale, 1.0), l
        dscale = float(jac.abs().max())
        assert float((df.double().cpu() - jac).abs().max()) < 5e-5 * max(dscale, 1.0), l


def test_filter_spline_matches_oracle(model_bundle):
    """The per-model quintic B-spline of the filter (what the default message kernels evaluate from
    shared memory) against the FP64 filter of the reference and its exact d-derivative.  The spline as a
    function is within 2e-7 / 1e-5 per Angstrom of the filter (tests/test_spline_host.py, FP64 evaluation);
    evaluated in FP32 on the device the stated bound is |f - spline| <= 2e-6 max(|f|, 1),
    |f' - spline'| <= 2e-5 max(|f'|, 1) -- ten times tighter than what the filter-table kernels are held to."""
    model, state, cfg = model_bundle
    eng = model.engine()
    w64 = po.to_torch_weights(state, torch.float64)
    rc = cfg["cutoff"]
    d = torch.cat([torch.linspace(0.05, rc, 6000), torch.tensor([rc, rc - 1e-6, 0.9572, 1e-3])]).float()
    for l in range(cfg["num_interactions"]):
        f, df = eng.filter_spline(l, d)
        p = f"interactions.{l}.message.rbf_to_scalar."

        def filt_fn(x):
            rbf = po.edge_features(w64, torch.stack([x, torch.zeros_like(x), torch.zeros_like(x)], 1), rc)[2]
            return torch.nn.functional.linear(
                torch.nn.functional.silu(torch.nn.functional.linear(rbf, w64[p + "0.weight"], w64[p + "0.bias"])),
                w64[p + "2.weight"], w64[p + "2.bias"])

        dd = d.double()
        filt, jac = torch.autograd.functional.jvp(filt_fn, dd, torch.ones_like(dd))
        scale = max(float(filt.abs().max()), 1.0)
        dscale = max(float(jac.abs().max()), 1.0)
        assert float((f.double().cpu() - filt).abs().max()) < 2e-6 * scale, l
        # at d == rc the reference's mask kills the derivative; the spline's one-sided derivative is ~1e-8
        assert float((df.double().cpu() - jac).abs().max()) < 2e-5 * dscale, l


def _oracle_stages(state, cfg, z, pos, off):
    w = po.to_torch_weights(state, torch.float64)
    e, f, keep = po.energy_and_forces_with_adjoints(
        w, torch.from_numpy(np.asarray(z, np.int64)), torch.from_numpy(np.asarray(pos, np.float32)).double(),
        cfg["cutoff"], po.batch_from_offsets(off))
    return e, f, keep


def test_stage_intermediates_and_adjoints(variant):
    _stage_check(variant, "fp32")


def test_stage_intermediates_and_adjoints_tensor_core():
    _stage_check("original", "tc")


def test_stage_intermediates_and_adjoints_filter_table(variant):
    _stage_check(variant, "fp32", "table")


def test_stage_intermediates_and_adjoints_filter_table_tensor_core():
    _stage_check("original", "tc", "table")


def test_stage_intermediates_and_adjoints_throughput_kernels(variant):
    """Same check with the small-system path switched off, so the ragged batch runs through the
    four-rows-per-warp spline kernels (pair-once layer 0, reverse-edge adjoints for the other layers)."""
    os.environ["MLFFD_SMALL_ROWS"] = "0"
    try:
        _stage_check(variant, "fp32", per_edge_adjoints=False)
    finally:
        os.environ.pop("MLFFD_SMALL_ROWS", None)


def _stage_check(variant, precision, filter_mode="spline", per_edge_adjoints=True):
    os.environ["MLFFD_DEBUG_KEEP"] = "1"
    try:
        model, state, cfg = _model(variant, precision=precision, filter_mode=filter_mode)
        gold = load_golden(variant)
        case = "ragged"
        z, pos, off = gold[f"{case}_numbers"], gold[f"{case}_positions"], gold[f"{case}_offsets"]
        e, f = _run(model, z, pos, off)
        eng = model.engine()
        e64, f64, keep = _oracle_stages(state, cfg, z, pos, off)
        H, L = cfg["hidden_dim"], cfg["num_interactions"]
        report = {}

        def check(name, got, ref, tol):
            got = got.double().cpu().reshape(ref.shape)
            err = float((got - ref).abs().max())
            scale = max(float(ref.abs().max()), 1.0)
            report[name] = {"err": err, "scale": scale, "ok": bool(err <= tol * scale)}

        # CSR entry k is (row = dst, col = src) at the lexicographic rank of (dst, src); the
        # oracle's edge k is (src, dst) at the rank of (src, dst).  So oracle edge k is the CUDA
        # edge rev[k]: permute every per-edge CUDA buffer by rev before comparing.
        rev = eng.debug_buffer("rev").long()
        geo = eng.debug_buffer("geo")[rev]
        check("unit", geo[:, :3], keep["unit"].detach(), 1e-6)
        check("dist", geo[:, 3], keep["d"].detach(), 1e-6)
        pair = eng.debug_buffer("pair").long()[rev]
        for l in range(L):
            if filter_mode == "table":
                filt = eng.debug_buffer("filter", l).reshape(-1, 3 * H)[pair]
            else:   # nothing per-pair is materialised: evaluate the spline the message kernels use
                filt = eng.filter_spline(l, geo[:, 3].contiguous())[0]
            ref = keep[f"filter{l}"].detach().clone()
            if l == 0:  # vector gate b is skipped for layer 0 (v_in == 0)
                filt = filt.clone(); filt[:, H:2 * H] = 0; ref[:, H:2 * H] = 0
            check(f"filter{l}", filt, ref, 2e-5)
            check(f"s_msg{l}", eng.debug_buffer("s_msg", l), keep[f"s_msg{l}"].detach(), 2e-5)
            check(f"v_msg{l}", eng.debug_buffer("v_msg", l), keep[f"v_msg{l}"].detach(), 2e-5)
            if precision == "fp32":  # the tensor-core reverse pass reuses y1 for its adjoint
                check(f"y1_{l}", eng.debug_buffer("y1", l), keep[f"y1_{l}"].detach(), 2e-5)
            if l < L - 1:
                g = eng.debug_buffer("gates", l).reshape(-1, 2 * H)
                check(f"g1_{l}", g[:, :H], keep[f"g1_{l}"].detach(), 2e-5)
                check(f"g2_{l}", g[:, H:], keep[f"g2_{l}"].detach(), 2e-5)
                check(f"s_in{l + 1}", eng.debug_buffer("s_in", l + 1), keep[f"s_in{l + 1}"].detach(), 2e-5)
                check(f"v_in{l + 1}", eng.debug_buffer("v_in", l + 1), keep[f"v_in{l + 1}"].detach(), 2e-5)
        check("s_out", eng.debug_buffer("s_out"), keep["s_out"].detach(), 2e-5)
        check("atom_energy", eng.debug_buffer("atom_energy"), keep["atomic_energies"].detach().reshape(-1), 2e-5)
        for l in range(L):
            check(f"sbar_msg{l}", eng.debug_buffer("sbar", l), keep[f"s_msg{l}"].grad, 5e-5)
            check(f"vbar_msg{l}", eng.debug_buffer("vbar", l), keep[f"v_msg{l}"].grad, 5e-5)
        if per_edge_adjoints:   # the pair-once layer-0 kernel only keeps u_bar_e - u_bar_rev(e): forces cover it
            adj = eng.debug_buffer("edge_adj")[rev]
            check("ubar", adj[:, :3], keep["unit"].grad, 5e-5)
        # d_bar through the filters only = dE/d(edge_rbf) . d(edge_rbf)/dd is not retained by the
        # oracle directly; forces cover it.
        check("forces", torch.from_numpy(f), f64, 2e-5)
        OUT.mkdir(exist_ok=True)
        (OUT / f"stages_{variant}_{precision}_{filter_mode}.json").write_text(json.dumps(report, indent=1))
        bad = {k: v for k, v in report.items() if not v["ok"]}
        assert not bad, bad
    finally:
        os.environ.pop("MLFFD_DEBUG_KEEP", None)


def test_energy_forces_match_golden(model_bundle, golden, variant):
    model, state, cfg = model_bundle
    report = {}
    worst_e = worst_f = 0.0
    for case in golden_cases(golden):
        z, pos, off = golden[f"{case}_numbers"], golden[f"{case}_positions"], golden[f"{case}_offsets"]
        e, f = _run(model, z, pos, off)
        natoms = np.diff(off)
        de32 = float(np.max(np.abs(e - golden[f"{case}_energy32"]) / natoms))
        de64 = float(np.max(np.abs(e - golden[f"{case}_energy64"]) / natoms))
        df32 = float(np.max(np.abs(f - golden[f"{case}_forces32"])))
        df64 = float(np.max(np.abs(f - golden[f"{case}_forces64"])))
        ref_noise = float(np.max(np.abs(golden[f"{case}_forces32"] - golden[f"{case}_forces64"])))
        report[case] = {"dE32_per_atom": de32, "dE64_per_atom": de64, "dF32": df32, "dF64": df64,
                        "ref_fp32_vs_fp64_F": ref_noise}
        assert de32 <= E_TOL and de64 <= E_TOL, (case, de32, de64)
        if case.endswith("_exact"):
            continue
        assert df64 <= max(F_TOL, 2 * ref_noise), (case, df64, ref_noise)
        assert df32 <= max(F_TOL, 2 * ref_noise), (case, df32, ref_noise)
        worst_e, worst_f = max(worst_e, de32), max(worst_f, df64)
    OUT.mkdir(exist_ok=True)
    (OUT / f"parity_{variant}.json").write_text(json.dumps(report, indent=1))


def test_batch_of_druglike_structures_matches_oracle(model_bundle):
    from mlff_distiller_b200 import synthetic
    model, state, cfg = model_bundle
    structs = synthetic.druglike_batch(48, first=100) + synthetic.druglike_batch(16, first=300, ragged=True)
    z, pos, off = synthetic.concatenate(structs)
    pos = pos.astype(np.float32)
    e, f = _run(model, z, pos, off)
    e_ref, f_ref = po.evaluate(state, cfg["cutoff"], z, pos, off, dtype=torch.float64, dense_graph=False)
    assert np.max(np.abs(e - e_ref) / np.diff(off)) <= E_TOL
    assert np.max(np.abs(f - f_ref)) <= F_TOL


def test_c2_full_size_batch_properties(model_bundle, variant):
    """BASELINE config C2 at its full size (1024 x 50 atoms) through size-independent properties, where the
    CPU oracle would take minutes: (1) the order of the structures in the batch is irrelevant -- bit for bit,
    every row is accumulated by one lane group in CSR order whatever CTA it lands in; (2) a structure has
    the same result in the full batch and in a half batch; (3) the forces of every structure sum to zero
    (translation invariance of the energy) up to FP32 rounding; (4) a spot sample of structures against the
    FP64 oracle."""
    from mlff_distiller_b200 import synthetic
    model, state, cfg = model_bundle
    structs = synthetic.druglike_batch(1024, first=0, n=50)
    z, pos, off = synthetic.concatenate(structs)
    pos = pos.astype(np.float32)
    e, f = _run(model, z, pos, off)
    assert np.isfinite(e).all() and np.isfinite(f).all()
    # (1) reversed structure order
    zr, posr, offr = synthetic.concatenate(structs[::-1])
    er, fr = _run(model, zr, posr.astype(np.float32), offr)
    assert np.array_equal(er[::-1], e)
    assert np.array_equal(fr.reshape(1024, 50, 3)[::-1].reshape(-1, 3), f)
    # (2) first half alone
    n_half = int(off[512])
    eh, fh = _run(model, z[:n_half], pos[:n_half], off[:513])
    assert np.array_equal(eh, e[:512]) and np.array_equal(fh, f[:n_half])
    # (3) net force per structure
    net = np.abs(f.reshape(1024, 50, 3).sum(axis=1)).max()
    assert net <= 2e-4, net
    # (4) spot sample against the oracle
    pick = [0, 511, 1023]
    zs, ps, os_ = synthetic.concatenate([structs[i] for i in pick])
    e_ref, f_ref = po.evaluate(state, cfg["cutoff"], zs, ps.astype(np.float32), os_, dtype=torch.float64,
                               dense_graph=False)
    for k, i in enumerate(pick):
        assert abs(e[i] - e_ref[k]) / 50 <= E_TOL
        assert np.max(np.abs(f[50 * i:50 * i + 50] - f_ref[50 * k:50 * k + 50])) <= F_TOL


def test_periodic_water_box_matches_oracle(model_bundle):
    from mlff_distiller_b200 import synthetic
    model, state, cfg = model_bundle
    box = synthetic.water_box(n_mol=64, seed=3001)   # L = 12.42 A >= 2 rc
    z, pos = box.numbers, box.positions.astype(np.float32)
    off = [0, len(z)]
    e, f = _run(model, z, pos, off, box.cell[None], box.pbc[None])
    e_ref, f_ref = po.evaluate(state, cfg["cutoff"], z, pos, off, box.cell[None], box.pbc[None],
                               dtype=torch.float64)
    assert abs(e[0] - e_ref[0]) / len(z) <= E_TOL
    fmax = float(np.abs(f_ref).max())
    assert np.max(np.abs(f - f_ref)) <= max(F_TOL, 2e-6 * fmax)


def test_reference_invariance_tests_on_its_water_fixture(model_bundle):
    """The reference's own physical-constraint tests at the reference's own tolerance
    (tests/unit/test_student_model.py:205-221 translation, :271-290 permutation: atol 1e-5 on the energy
    of its water fixture, :71-77)."""
    model, state, cfg = model_bundle
    z = np.array([8, 1, 1])
    pos = np.float32([[0, 0, 0], [0.96, 0, 0], [-0.24, 0.93, 0]])
    e0, f0 = _run(model, z, pos, [0, 3])
    shift = (np.random.default_rng(0).normal(size=3) * 10.0).astype(np.float32)
    e1, f1 = _run(model, z, pos + shift, [0, 3])
    assert abs(e1[0] - e0[0]) <= 1e-5, (e0, e1)
    perm = np.array([2, 0, 1])
    e2, f2 = _run(model, z[perm], pos[perm], [0, 3])
    assert abs(e2[0] - e0[0]) <= 1e-5 and np.abs(f2 - f0[perm]).max() <= 1e-5


def test_invariances_and_extensivity(model_bundle):
    """Same properties on a 40-atom structure (|E| ~ 250 eV, where one FP32 ulp of the energy is already
    1.5e-5 eV).  A translation changes the FP32 rounding of every coordinate, so the bound is the FP32
    bound of the north-star per atom (1e-5 eV/atom, 1e-4 eV/A) or twice what the reference's own FP32
    arithmetic (the CPU oracle) moves under the same operation, whichever is larger."""
    from mlff_distiller_b200 import synthetic
    model, state, cfg = model_bundle
    s = synthetic.druglike(31337, 40)
    z, pos = s.numbers, s.positions.astype(np.float32)
    shift = np.float32([3.0, -2.0, 1.5])
    e0, f0 = _run(model, z, pos, [0, 40])
    r0 = po.evaluate(state, cfg["cutoff"], z, pos, [0, 40])
    r1 = po.evaluate(state, cfg["cutoff"], z, pos + shift, [0, 40])
    e_tol = max(40 * E_TOL, 2 * abs(float(r1[0][0] - r0[0][0])))
    f_tol = max(F_TOL, 2 * float(np.abs(r1[1] - r0[1]).max()))
    e1, f1 = _run(model, z, pos + shift, [0, 40])
    assert abs(e1[0] - e0[0]) <= e_tol and np.abs(f1 - f0).max() <= f_tol
    # permutation
    perm = np.random.default_rng(0).permutation(40)
    e2, f2 = _run(model, z[perm], pos[perm], [0, 40])
    assert abs(e2[0] - e0[0]) <= e_tol and np.abs(f2 - f0[perm]).max() <= f_tol
    # extensivity: two copies 30 A apart, as one structure and as a batch of two
    pos2 = np.concatenate([pos, pos + np.float32([30.0, 0, 0])])
    e3, f3 = _run(model, np.concatenate([z, z]), pos2, [0, 80])
    assert abs(e3[0] - 2 * e0[0]) <= 2 * e_tol
    e4, f4 = _run(model, np.concatenate([z, z]), pos2, [0, 40, 80])
    assert np.abs(e4 - e0[0]).max() <= e_tol and np.abs(f4 - np.concatenate([f0, f0])).max() <= f_tol
    assert np.abs(f0.sum(0)).max() < 1e-4  # Newton III


def test_periodic_water_box_original_weights_3000_atoms():
    """VERDICT weak #3: the 427K Original model on a periodic box of 3 000 atoms (cell list + minimum
    image, tensor-core update block, spline filter) against the FP64 oracle on the same edges."""
    from mlff_distiller_b200 import synthetic
    model, state, cfg = _model("original", precision="tc", pbc_mode="minimum_image")
    box = synthetic.water_box(n_mol=1000, seed=3003)
    z, pos = box.numbers, box.positions.astype(np.float32)
    off = [0, len(z)]
    e, f = _run(model, z, pos, off, box.cell[None], box.pbc[None])
    assert model.engine().status().num_edges > 150_000
    e_ref, f_ref = po.evaluate(state, cfg["cutoff"], z, pos, off, box.cell[None], box.pbc[None],
                               dtype=torch.float64)
    assert abs(e[0] - e_ref[0]) / len(z) <= E_TOL
    assert np.max(np.abs(f - f_ref)) <= max(F_TOL, 2e-6 * float(np.abs(f_ref).max()))


def test_forces_are_the_energy_gradient(model_bundle):
    """Central finite differences of the CUDA energy (FP32 energies: loose bound as in the
    reference's own test, tests/unit/test_student_model.py:377-411: eps 1e-4... max err 5e-3)."""
    from mlff_distiller_b200 import synthetic
    model, state, cfg = model_bundle
    s = synthetic.druglike(99, 12)
    z, pos = s.numbers, s.positions.astype(np.float32)
    _, f = _run(model, z, pos, [0, 12])
    h = 2e-2
    rng = np.random.default_rng(1)
    for _ in range(6):
        a, k = int(rng.integers(12)), int(rng.integers(3))
        pp, pm = pos.copy(), pos.copy()
        pp[a, k] += h
        pm[a, k] -= h
        ep, _ = _run(model, z, pp, [0, 12])
        em, _ = _run(model, z, pm, [0, 12])
        fd = -(ep[0] - em[0]) / (float(pp[a, k]) - float(pm[a, k]))
        assert abs(fd - f[a, k]) < 2e-2 + 2e-2 * abs(f[a, k])


def test_deterministic_and_autograd_drop_in(model_bundle, golden):
    model, state, cfg = model_bundle
    z, pos, off = golden["batch4x50_numbers"], golden["batch4x50_positions"], golden["batch4x50_offsets"]
    e1, f1 = _run(model, z, pos, off)
    e2, f2 = _run(model, z, pos, off)
    assert np.array_equal(e1, e2) and np.array_equal(f1, f2)
    # the reference recipe: energies = model(Z, R, batch=...); F = -autograd.grad(...)
    zt = torch.from_numpy(z).cuda()
    pt = torch.from_numpy(pos).cuda().requires_grad_(True)
    batch = po.batch_from_offsets(off).cuda()
    energies = model(zt, pt, cell=None, pbc=None, batch=batch)
    assert energies.shape == (4,)
    forces = -torch.autograd.grad(energies, pt, grad_outputs=torch.ones_like(energies))[0]
    assert np.array_equal(forces.double().cpu().numpy(), f1)
    # single structure: 0-dim energy, predict_energy_and_forces
    zs = torch.from_numpy(golden["h2o_numbers"]).cuda()
    ps = torch.from_numpy(golden["h2o_positions"]).cuda()
    e, f = model.predict_energy_and_forces(zs, ps)
    assert e.dim() == 0 and f.shape == (3, 3)
    e_only = model(zs, ps)
    assert e_only.dim() == 0 and abs(float(e_only) - float(e)) < 1e-6


def test_edge_capacity_overflow_recovers(variant):
    model, state, cfg = _model(variant)
    eng = model.engine()
    eng.reserve(64, 8, 4)  # far too few edges on purpose
    gold = load_golden(variant)
    z, pos, off = gold["drug50_numbers"], gold["drug50_positions"], gold["drug50_offsets"]
    e, f = _run(model, z, pos, off)
    assert abs(e[0] - gold["drug50_energy32"][0]) / 50 <= E_TOL


# ---------------------------------------------------------------------------------------------
# tensor-core path (tcgen05, two-term FP16 split): must meet the SAME FP32 bounds
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant_name", ["original", "tiny", "ultra_tiny"])
def test_tensor_core_filter_table_matches_ffma_and_oracle(variant_name):
    model_tc, state, cfg = _model(variant_name, precision="tc")
    model_32, _, _ = _model(variant_name, precision="fp32")
    rc = cfg["cutoff"]
    d = torch.cat([torch.linspace(0.4, rc + 0.3, 4099), torch.tensor([rc, rc - 1e-6, 0.9572])]).float()
    for l in range(cfg["num_interactions"]):
        f_tc, df_tc = model_tc.engine().filter_table(l, d)
        f_32, df_32 = model_32.engine().filter_table(l, d)
        scale = max(float(f_32.abs().max()), 1.0)
        dscale = max(float(df_32.abs().max()), 1.0)
        assert float((f_tc - f_32).abs().max()) < 4e-6 * scale, l
        assert float((df_tc - df_32).abs().max()) < 4e-6 * dscale, l


def test_tensor_core_energy_forces_match_golden():
    model, state, cfg = _model("original", precision="tc")
    gold = load_golden("original")
    report = {}
    for case in golden_cases(gold):
        z, pos, off = gold[f"{case}_numbers"], gold[f"{case}_positions"], gold[f"{case}_offsets"]
        e, f = _run(model, z, pos, off)
        natoms = np.diff(off)
        de64 = float(np.max(np.abs(e - gold[f"{case}_energy64"]) / natoms))
        df64 = float(np.max(np.abs(f - gold[f"{case}_forces64"])))
        ref_noise = float(np.max(np.abs(gold[f"{case}_forces32"] - gold[f"{case}_forces64"])))
        report[case] = {"dE64_per_atom": de64, "dF64": df64, "ref_fp32_vs_fp64_F": ref_noise}
        assert de64 <= E_TOL, (case, de64)
        if not case.endswith("_exact"):
            assert df64 <= max(F_TOL, 2 * ref_noise), (case, df64)
    OUT.mkdir(exist_ok=True)
    (OUT / "parity_original_tc.json").write_text(json.dumps(report, indent=1))


def test_tensor_core_batch_matches_oracle():
    from mlff_distiller_b200 import synthetic
    model, state, cfg = _model("original", precision="tc")
    structs = synthetic.druglike_batch(130, first=700)   # 6500 atoms: several tiles per CTA incl. a ragged tail
    z, pos, off = synthetic.concatenate(structs)
    pos = pos.astype(np.float32)
    e, f = _run(model, z, pos, off)
    e2, f2 = _run(model, z, pos, off)
    assert np.array_equal(e, e2) and np.array_equal(f, f2)
    e_ref, f_ref = po.evaluate(state, cfg["cutoff"], z, pos, off, dtype=torch.float64, dense_graph=False)
    assert np.max(np.abs(e - e_ref) / np.diff(off)) <= E_TOL
    assert np.max(np.abs(f - f_ref)) <= F_TOL


# ---------------------------------------------------------------------------------------------
# cell-list candidate generator (large structures): same edges, bit for bit
# ---------------------------------------------------------------------------------------------
def _with_neighbor_mode(mode, variant="ultra_tiny"):
    os.environ["MLFFD_NEIGHBOR"] = mode
    try:
        bundle = _model(variant)
        bundle[0].engine()   # the context reads MLFFD_NEIGHBOR when it is created
        return bundle
    finally:
        os.environ.pop("MLFFD_NEIGHBOR", None)


def test_cell_list_edges_bit_exact_small_and_batched():
    from mlff_distiller_b200.student_model import radius_graph
    model, state, cfg = _with_neighbor_mode("cells")
    gold = load_golden("ultra_tiny")
    for case in golden_cases(gold):
        pos, off = gold[f"{case}_positions"], gold[f"{case}_offsets"]
        ei = radius_graph(torch.from_numpy(pos), cfg["cutoff"], po.batch_from_offsets(off).cuda(),
                          engine=model.engine()).cpu().numpy()
        assert np.array_equal(ei, gold[f"{case}_edge_index"]), case
    rng = np.random.default_rng(21)
    n, L = 700, 16.0
    pos = rng.uniform(-5.0, L + 5.0, size=(n, 3)).astype(np.float32)
    cell = np.array([[L, 0, 0], [2.0, L, 0], [1.0, -1.5, L + 3]], dtype=np.float64)
    for pbc in ([True, True, True], [True, False, True], [False, False, True]):
        ei_ref, _ = po.neighbor_list(pos, [0, n], cfg["cutoff"], cell[None], np.array([pbc]))
        ei = radius_graph(torch.from_numpy(pos), cfg["cutoff"], None, engine=model.engine(),
                          cell=torch.from_numpy(cell), pbc=torch.tensor(pbc)).cpu().numpy()
        assert np.array_equal(ei, ei_ref), pbc


def test_cell_list_water_box_10k_atoms():
    """BASELINE config C4: 9 999-atom periodic water box, automatic path selection (cells)."""
    from mlff_distiller_b200 import synthetic
    from mlff_distiller_b200.student_model import radius_graph
    model, state, cfg = _model("ultra_tiny")
    box = synthetic.water_box()
    pos = box.positions.astype(np.float32)
    ei = radius_graph(torch.from_numpy(pos), cfg["cutoff"], None, engine=model.engine(),
                      cell=torch.from_numpy(box.cell), pbc=torch.from_numpy(box.pbc)).cpu().numpy()
    ei_ref, _ = po.neighbor_list(pos, [0, len(pos)], cfg["cutoff"], box.cell[None], box.pbc[None])
    assert ei.shape[1] > 400_000
    assert np.array_equal(ei, ei_ref)
    # open-boundary blob of the same size: cells vs the reference semantics
    blob = (np.random.default_rng(3).uniform(0, 60.0, size=(6000, 3))).astype(np.float32)
    ei2 = radius_graph(torch.from_numpy(blob), cfg["cutoff"], None, engine=model.engine()).cpu().numpy()
    ei2_ref, _ = po.neighbor_list(blob, [0, 6000], cfg["cutoff"])
    assert np.array_equal(ei2, ei2_ref)


def test_cell_list_energy_forces_periodic_1500_atoms():
    from mlff_distiller_b200 import synthetic
    model, state, cfg = _with_neighbor_mode("cells", "tiny")
    box = synthetic.water_box(n_mol=500, seed=3002)
    z, pos = box.numbers, box.positions.astype(np.float32)
    off = [0, len(z)]
    e, f = _run(model, z, pos, off, box.cell[None], box.pbc[None])
    e_ref, f_ref = po.evaluate(state, cfg["cutoff"], z, pos, off, box.cell[None], box.pbc[None],
                               dtype=torch.float64)
    assert abs(e[0] - e_ref[0]) / len(z) <= E_TOL
    assert np.max(np.abs(f - f_ref)) <= max(F_TOL, 2e-6 * float(np.abs(f_ref).max()))


@pytest.mark.parametrize("variant_name", ["original", "tiny", "ultra_tiny"])
def test_message_kernel_variants_agree(variant_name):
    """cp.async-pipelined message kernels == plain kernels bit for bit; the pair-once reverse pass
    differs from the per-directed-edge one only by rounding."""
    from mlff_distiller_b200 import synthetic
    structs = (synthetic.druglike_batch(40, first=300, ragged=True) + [synthetic.water(), synthetic.Structure([6], [[0, 0, 0]])]
               + [synthetic.alkane_chain(40)])
    z, pos, off = synthetic.concatenate(structs)
    dev = "cuda:0"
    z_d = torch.from_numpy(z.astype(np.int32)).to(dev)
    p_d = torch.from_numpy(pos.astype(np.float32)).to(dev)
    o_d = torch.from_numpy(off.astype(np.int32)).to(dev)

    def run(env):
        env = {"MLFFD_MSG_TEAM": "0", **env}   # row-per-warp kernels whatever the batch size
        os.environ.update(env)
        try:
            model, state, cfg = _model(variant_name, filter_mode="table")
            model.engine()
        finally:
            for k in env:
                os.environ.pop(k, None)
        e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(structs))
        return e.cpu().numpy(), f.cpu().numpy()

    e_def, f_def = run({"MLFFD_MSG_FWD": "pipe", "MLFFD_MSG_BWD": "pipe"})   # falls back to pairs for H != 128
    e_pairs, f_pairs = run({"MLFFD_MSG_FWD": "rows", "MLFFD_MSG_BWD": "pairs"})
    assert np.array_equal(e_def, e_pairs) and np.array_equal(f_def, f_pairs)
    for depth in ("2", "8"):
        e_d, f_d = run({"MLFFD_MSG_FWD": "pipe", "MLFFD_MSG_BWD": "pipe", "MLFFD_PIPE_DEPTH_FWD": depth,
                        "MLFFD_PIPE_DEPTH_BWD": "2" if depth == "2" else "4"})
        assert np.array_equal(e_def, e_d) and np.array_equal(f_def, f_d)
    e_edges, f_edges = run({"MLFFD_MSG_FWD": "rows", "MLFFD_MSG_BWD": "edges"})
    assert np.array_equal(e_def, e_edges)
    assert np.max(np.abs(f_def - f_edges)) <= 2e-5


def test_tiled_readout_matches_warp_readout_and_oracle():
    """readout_tile_kernel (H = 128, large batches) against the warp kernel and the FP64 oracle."""
    from mlff_distiller_b200 import synthetic
    structs = synthetic.druglike_batch(9, first=700, ragged=True) + [synthetic.water()]
    z, pos, off = synthetic.concatenate(structs)
    pos32 = pos.astype(np.float32)
    dev = "cuda:0"
    z_d = torch.from_numpy(z.astype(np.int32)).to(dev)
    p_d = torch.from_numpy(pos32).to(dev)
    o_d = torch.from_numpy(off.astype(np.int32)).to(dev)
    out = {}
    for mode in ("tile", "warp"):
        os.environ["MLFFD_READOUT"] = mode
        try:
            model, state, cfg = _model("original")
            model.engine()
        finally:
            os.environ.pop("MLFFD_READOUT", None)
        e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(structs))
        e0 = model.forward(torch.from_numpy(z), torch.from_numpy(pos32), batch=po.batch_from_offsets(off))
        assert np.array_equal(e0.cpu().numpy(), e.cpu().numpy())      # energy-only path, same kernel
        out[mode] = (e.cpu().numpy().astype(np.float64), f.cpu().numpy().astype(np.float64))
    e_ref, f_ref = po.evaluate(state, cfg["cutoff"], z, pos32.astype(np.float64), off, dtype=torch.float64)
    counts = np.diff(off)
    for mode in out:
        assert np.max(np.abs(out[mode][0] - e_ref) / counts) <= E_TOL, mode
        assert np.max(np.abs(out[mode][1] - f_ref)) <= 3e-5, mode
    assert np.max(np.abs(out["tile"][0] - out["warp"][0]) / counts) <= 2e-6
    assert np.max(np.abs(out["tile"][1] - out["warp"][1])) <= 1e-5


# ---------------------------------------------------------------------------------------------
# single-pass tensor-core modes: looser, STATED bounds
# ---------------------------------------------------------------------------------------------
LOOSE_BOUNDS = {            # (eV / atom, eV / A) against the reference's FP64 outputs
    "tc_fp16": (2e-3, 5e-2),   # one product of FP16-rounded operands (11-bit significands, TF32 class)
    "tc_bf16": (2e-2, 5e-1),   # one product of BF16-rounded operands (8-bit significands)
}


@pytest.mark.parametrize("mode", sorted(LOOSE_BOUNDS))
@pytest.mark.parametrize("variant_name", ["original", "tiny"])
def test_single_pass_tensor_core_modes_within_stated_bounds(variant_name, mode):
    e_tol, f_tol = LOOSE_BOUNDS[mode]
    # the single-pass products act on the dense filter layers: those run per step in table mode only
    model, state, cfg = _model(variant_name, precision=mode, filter_mode="table")
    gold = load_golden(variant_name)
    worst_e = worst_f = 0.0
    for case in golden_cases(gold):
        z, pos, off = gold[f"{case}_numbers"], gold[f"{case}_positions"], gold[f"{case}_offsets"]
        e, f = _run(model, z, pos, off)
        e2, f2 = _run(model, z, pos, off)
        assert np.array_equal(e, e2) and np.array_equal(f, f2)
        worst_e = max(worst_e, float(np.max(np.abs(e - gold[f"{case}_energy64"]) / np.diff(off))))
        if not case.endswith("_exact"):
            worst_f = max(worst_f, float(np.max(np.abs(f - gold[f"{case}_forces64"]))))
    assert worst_e <= e_tol and worst_f <= f_tol, (worst_e, worst_f)
    # and they really are a different arithmetic: coarser than the FP32-equivalent default
    assert worst_e > E_TOL or worst_f > F_TOL


# ---------------------------------------------------------------------------------------------
# small-system latency path (single-trajectory MD) vs the throughput path
# ---------------------------------------------------------------------------------------------
def _with_env(env, variant="original", **kw):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        bundle = _model(variant, **kw)
        bundle[0].engine()   # the context reads its tuning variables when it is created
        return bundle
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_small_system_kernels_match_throughput_kernels_and_golden():
    """N <= 2048 atoms takes the latency path (one-launch neighbour list for N <= 64, all layers'
    filter tables in one launch, many-block FFMA update block, block-per-4-atoms read-out);
    MLFFD_SMALL_ROWS=0 forces the throughput kernels (tcgen05 update block, warp read-out,
    multi-kernel neighbour list).  Both must meet the FP32 bounds on every golden case, produce the
    same edges bit for bit, and agree with each other far inside the tolerance."""
    small, _, _ = _model("original", precision="tc")
    big, _, _ = _with_env({"MLFFD_SMALL_ROWS": "0", "MLFFD_FILTER_BATCH": "0"}, precision="tc")
    gold = load_golden("original")
    for case in golden_cases(gold):
        z, pos, off = gold[f"{case}_numbers"], gold[f"{case}_positions"], gold[f"{case}_offsets"]
        e_s, f_s = _run(small, z, pos, off)
        edges_s = small.engine().export_edges().cpu().numpy()
        e_b, f_b = _run(big, z, pos, off)
        edges_b = big.engine().export_edges().cpu().numpy()
        assert np.array_equal(edges_s, edges_b), case
        natoms = np.diff(off)
        for e, f in ((e_s, f_s), (e_b, f_b)):
            assert np.max(np.abs(e - gold[f"{case}_energy64"]) / natoms) <= E_TOL, case
            if not case.endswith("_exact"):
                ref_noise = float(np.max(np.abs(gold[f"{case}_forces32"] - gold[f"{case}_forces64"])))
                assert np.max(np.abs(f - gold[f"{case}_forces64"])) <= max(F_TOL, 2 * ref_noise), case
        assert np.max(np.abs(e_s - e_b) / natoms) <= 5e-6, case   # single atom: 2.9e-6 (split-FP16 rounding of the tensor-core path)
        if not case.endswith("_exact"):
            assert np.max(np.abs(f_s - f_b)) <= 4e-5, case


def test_small_neighbor_kernel_periodic_and_batched_bit_exact():
    """The one-launch neighbour kernel (N <= 64) against the multi-kernel sweep: periodic cell and
    a batch of tiny structures, every index array identical."""
    from mlff_distiller_b200 import synthetic
    from mlff_distiller_b200.student_model import StudentForceField
    small, _, _ = _model("ultra_tiny")
    sweep, _, _ = _with_env({"MLFFD_NEIGHBOR": "sweep"}, "ultra_tiny")
    rng = np.random.default_rng(5)
    structs = synthetic.druglike_batch(3, first=40, n=16)
    z, pos, off = synthetic.concatenate(structs)
    cell = np.array([[11.0, 0.0, 0.0], [1.5, 10.5, 0.0], [0.5, -1.0, 12.0]])
    cases = [(z, pos.astype(np.float32), off, None, None),
             (z[:40], (pos[:40] + rng.normal(0, 0.2, (40, 3))).astype(np.float32), np.array([0, 40]), cell[None], np.array([[True, True, False]]))]
    for zc, pc, oc, cells, pbc in cases:
        out = []
        for model in (small, sweep):
            model.pbc_mode = "minimum_image"
            p_d = torch.from_numpy(pc).cuda()
            o_d = torch.from_numpy(np.asarray(oc, dtype=np.int32)).cuda()
            c_d = b_d = None
            if cells is not None:
                c_d, b_d = StudentForceField.pack_cells(torch.from_numpy(cells), torch.from_numpy(pbc), len(oc) - 1, "cuda:0")
            eng = model.engine()
            eng.ensure(len(zc), len(oc) - 1)
            eng.neighbor_list_async(p_d, o_d, len(oc) - 1, c_d, b_d)
            out.append({k: eng.debug_buffer(k).cpu().numpy() for k in ("rowptr", "col", "rev", "pair", "edge_dst", "geo", "pair_dist")})
        assert out[0]["col"].size > 0
        for k in out[0]:
            assert np.array_equal(out[0][k], out[1][k]), k


@pytest.mark.parametrize("variant_name", ["original", "tiny", "ultra_tiny"])
def test_team_message_kernels_match_row_per_warp_kernels(variant_name):
    """Small systems: a team of four warps shares every CSR row (message_team.cuh).  Same arithmetic
    as the row-per-warp kernels up to the order of the four partial sums; deterministic."""
    from mlff_distiller_b200 import synthetic
    structs = (synthetic.druglike_batch(12, first=640, ragged=True) + [synthetic.water(), synthetic.Structure([6], [[0, 0, 0]])]
               + [synthetic.alkane_chain(60)])
    z, pos, off = synthetic.concatenate(structs)
    # rattled: on the exact mirror-symmetric chain the vector features cancel to rounding noise and
    # d|v|/dv is arbitrary, so any change of summation order moves the forces
    pos = (pos + np.random.default_rng(21).normal(0.0, 0.05, pos.shape)).astype(np.float32)
    team, state, cfg = _with_env({"MLFFD_MSG_TEAM": "4"}, variant_name, filter_mode="table")
    rows, _, _ = _with_env({"MLFFD_MSG_TEAM": "0"}, variant_name, filter_mode="table")
    e_t, f_t = _run(team, z, pos, off)
    e_t2, f_t2 = _run(team, z, pos, off)
    e_r, f_r = _run(rows, z, pos, off)
    assert np.array_equal(e_t, e_t2) and np.array_equal(f_t, f_t2)
    assert not np.array_equal(f_t, f_r) or variant_name != "original"   # really a different kernel
    assert np.max(np.abs(e_t - e_r) / np.diff(off)) <= 2e-6
    assert np.max(np.abs(f_t - f_r)) <= 2e-5
    e_ref, f_ref = po.evaluate(state, cfg["cutoff"], z, pos, off, dtype=torch.float64, dense_graph=False)
    assert np.max(np.abs(e_t - e_ref) / np.diff(off)) <= E_TOL
    assert np.max(np.abs(f_t - f_ref)) <= F_TOL


# ---------------------------------------------------------------------------------------------
# Verlet-skin neighbour list: an option, bit-identical to the full rebuild
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("system", ["chain300_sweep", "water_box_3000_cells"])
def test_skin_list_is_bit_identical_to_the_full_rebuild(system):
    """A random walk of the positions: the exact list derived from the skin candidates must equal the
    full rebuild's in every array (row starts, sources, reverse indices, geometry), whether the step
    reuses the candidate list or rebuilds it; rebuilds must happen, and not on every step."""
    from mlff_distiller_b200 import synthetic
    from mlff_distiller_b200.student_model import StudentForceField
    exact, state, cfg = _model("ultra_tiny", pbc_mode="minimum_image")
    skinned, _, _ = _model("ultra_tiny", pbc_mode="minimum_image", skin=1.0)
    if system == "chain300_sweep":
        s = synthetic.alkane_chain(100)
        pos, cells, pbc = s.positions.astype(np.float32), None, None
    else:
        box = synthetic.water_box(n_mol=1000, seed=3004)
        pos, cells, pbc = box.positions.astype(np.float32), box.cell[None], box.pbc[None]
    n = len(pos)
    off = torch.tensor([0, n], dtype=torch.int32, device="cuda")
    c_d = b_d = None
    if cells is not None:
        c_d, b_d = StudentForceField.pack_cells(torch.from_numpy(cells), torch.from_numpy(pbc), 1, "cuda:0")
    rng = np.random.default_rng(5)
    steps = 40
    for step in range(steps):
        pos = (pos + rng.normal(0.0, 0.04, pos.shape)).astype(np.float32)
        p_d = torch.from_numpy(pos).cuda()
        out = []
        for model in (exact, skinned):
            eng = model.engine()
            eng.ensure(n, 1, 80)
            eng.neighbor_list_async(p_d, off, 1, c_d, b_d)
            st = eng.status()
            assert not st.overflow
            out.append({k: eng.debug_buffer(k).cpu().numpy() for k in ("rowptr", "col", "rev", "edge_dst", "geo")})
        assert out[0]["col"].size > 1000
        for k in out[0]:
            assert np.array_equal(out[0][k], out[1][k]), (step, k)
    rebuilds = int(skinned.engine().status().skin_rebuilds)
    assert 1 <= rebuilds < steps, rebuilds
    assert int(exact.engine().status().skin_rebuilds) == 0
    # energies and forces through the same lists
    z = np.full(n, 6, dtype=np.int64) if system == "chain300_sweep" else box.numbers
    e0, f0 = _run(exact, z, pos, [0, n], cells, pbc)
    e1, f1 = _run(skinned, z, pos, [0, n], cells, pbc)
    assert np.array_equal(e0, e1) and np.array_equal(f0, f1)
