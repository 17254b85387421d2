"""Calculator semantics on the GPU, mirroring the reference's
tests/integration/test_ase_calculator.py (which needs ASE + a trained checkpoint and cannot run
here): result types, n_calls, ASE-style caching, validation errors, calculate_batch, timing,
PBC handling, reset/repr, short NVE runs."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from mlff_distiller_b200 import md, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def calc():
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    return StudentForceFieldCalculator(GOLDEN / "weights_original.npz", device="cuda", enable_timing=True,
                                       use_compile=True, use_fp16=True, use_torch_cluster=False,
                                       use_analytical_forces=True, batch_size=8)  # reference kwargs accepted


def test_energy_forces_types_and_golden(calc):
    gold = load_golden("original")
    atoms = synthetic.water()
    atoms.calc = calc
    n0 = calc.n_calls
    e = atoms.get_potential_energy()
    f = atoms.get_forces()
    assert isinstance(e, float) and np.isfinite(e)
    assert f.shape == (3, 3) and f.dtype == np.float32 and np.isfinite(f).all()
    assert abs(e - float(gold["h2o_energy32"][0])) < 3e-5
    assert np.abs(f - gold["h2o_forces32"]).max() < 1e-4
    assert calc.n_calls == n0 + 1  # the forces came from the cache


def test_cache_and_n_calls(calc):
    atoms = synthetic.benzene()
    atoms.calc = calc
    n0 = calc.n_calls
    atoms.get_potential_energy(); atoms.get_forces(); atoms.get_potential_energy()
    assert calc.n_calls == n0 + 1
    atoms.set_positions(atoms.get_positions() + 0.01)
    atoms.get_forces()
    assert calc.n_calls == n0 + 2
    stats = calc.get_timing_stats()
    for key in ("n_calls", "total_time", "avg_time", "min_time", "max_time", "median_time"):
        assert key in stats
    assert calc.avg_time > 0


def test_validation_errors(calc):
    with pytest.raises(ValueError, match="empty structure"):
        calc.calculate(synthetic.Structure([], np.zeros((0, 3))))
    bad = synthetic.water()
    bad.positions[0, 0] = np.nan
    with pytest.raises(ValueError, match="NaN"):
        calc.calculate(bad)
    with pytest.raises(ValueError, match="Invalid atomic numbers"):
        calc.calculate(synthetic.Structure([0, 1], [[0, 0, 0], [1, 0, 0]]))
    with pytest.raises(ValueError, match="Invalid atomic numbers"):
        calc.calculate(synthetic.Structure([105], [[0, 0, 0]]))  # beyond the trained embedding (max_z = 100)


def test_calculate_batch_matches_single(calc):
    assert calc.calculate_batch([]) == []
    structs = [synthetic.water(), synthetic.benzene()] + synthetic.druglike_batch(3, first=50, ragged=True)
    res = calc.calculate_batch(structs)
    assert len(res) == len(structs)
    for s, r in zip(structs, res):
        assert isinstance(r["energy"], float) and r["forces"].shape == (len(s), 3)
        calc.calculate(s.copy())
        assert abs(r["energy"] - calc.results["energy"]) < 1e-4 * max(1.0, abs(r["energy"]) * 1e-2)
        assert np.abs(r["forces"] - calc.results["forces"]).max() < 1e-4
    one = calc.calculate_batch([synthetic.water()])
    assert len(one) == 1 and one[0]["forces"].shape == (3, 3) and one[0]["stress"] is None
    only_e = calc.calculate_batch(structs[:2], properties=["energy"])
    assert "forces" not in only_e[0]


def test_pbc_inputs(calc):
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    # reference behaviour (tests/integration/test_ase_calculator.py:255-278): PBC given, ignored
    cu = synthetic.Structure([29], [[0, 0, 0]], cell=np.array([[0, 1.79, 1.79], [1.79, 0, 1.79], [1.79, 1.79, 0]]),
                             pbc=[True, True, True])
    calc.calculate(cu)
    assert np.isfinite(calc.results["energy"]) and np.isfinite(calc.results["forces"]).all()
    pbc_calc = StudentForceFieldCalculator(GOLDEN / "weights_ultra_tiny.npz", device="cuda", pbc_mode="minimum_image")
    with pytest.raises(ValueError, match="2\\*cutoff"):
        pbc_calc.calculate(cu)
    box = synthetic.water_box(n_mol=64, seed=5)
    pbc_calc.calculate(box)
    e_pbc = pbc_calc.results["energy"]
    ign = StudentForceFieldCalculator(GOLDEN / "weights_ultra_tiny.npz", device="cuda")
    ign.calculate(box)
    assert abs(e_pbc - ign.results["energy"]) > 1e-3  # periodic images do contribute
    assert pbc_calc.implemented_properties == ["energy", "forces"]
    assert "stress" in StudentForceFieldCalculator(GOLDEN / "weights_ultra_tiny.npz", device="cuda",
                                                    enable_stress=True).implemented_properties


def test_reset_and_repr(calc):
    calc.calculate(synthetic.water())
    calc.reset()
    assert calc.results == {}
    assert "StudentForceFieldCalculator" in repr(calc) and "weights_original.npz" in repr(calc)


def test_descent_lowers_energy_and_short_nve(calc):
    atoms = synthetic.druglike(4321, 24)
    x = atoms.get_positions()
    work = atoms.copy()

    def ef(pos):
        work.set_positions(pos)
        calc.calculate(work)
        return calc.results["energy"], calc.results["forces"].astype(np.float64)

    e0, f = ef(x)
    for _ in range(20):
        x = x + 0.002 * f
        e1, f = ef(x)
    assert e1 < e0
    # reference test: 10-step NVE drift < 10 %
    m = atoms.get_masses()
    v0 = md.maxwell_boltzmann(m, 300.0, np.random.default_rng(0))
    out = md.velocity_verlet(ef, x, v0, m, steps=10, dt_fs=0.5)
    assert abs(out["drift_percent"]) < 10.0


def _nve_drift(calc, atoms, steps, seed=42):
    work = atoms.copy()

    def ef(pos):
        work.set_positions(pos)
        calc.calculate(work)
        return calc.results["energy"], calc.results["forces"]

    m = atoms.get_masses()
    v0 = md.maxwell_boltzmann(m, 300.0, np.random.default_rng(seed), atoms.get_positions(), zero_rotation=True)
    return md.velocity_verlet(ef, atoms.get_positions(), v0, m, steps=steps, dt_fs=0.5)


def test_nve_1000_steps_drift_within_reference_bar(calc):
    """BASELINE config C1: 1000 velocity-Verlet steps at 0.5 fs, T0 = 300 K, seed 42, driven through
    the calculator.  The reference's bar, |drift| <= 0.14 % (its README.md:82), comes from 12 - 32-atom
    organic molecules: asserted strictly on benzene.  For H2O the end-point drift is a noisy statistic
    of the three-atom trajectory itself (a 1e-6 A rattle of the start moves the 10 ps
    value between -0.5 and +0.1 %, whatever the kernels); observed 0.03 - 0.09 % over these 1000 steps,
    asserted at 0.30 % together with the amplitude of the energy oscillation."""
    out = _nve_drift(calc, synthetic.benzene(), 1000)
    assert abs(out["drift_percent"]) <= 0.14
    out = _nve_drift(calc, synthetic.water(), 1000)
    tot = np.asarray(out["total"])
    band = 100.0 * float(np.max(np.abs(tot - tot[0]))) / abs(tot[0])
    assert band <= 0.5, band                          # the oscillation itself stays small
    assert abs(out["drift_percent"]) <= 0.30, out["drift_percent"]


def test_device_md_matches_host_verlet(calc):
    """On-device velocity Verlet (graph replay) against the host float64 integrator driving the
    calculator: same initial state, same energy series."""
    atoms = synthetic.benzene()
    m = atoms.get_masses()
    v0 = md.maxwell_boltzmann(m, 300.0, np.random.default_rng(7), atoms.get_positions(), zero_rotation=True)
    work = atoms.copy()

    def ef(pos):
        work.set_positions(pos)
        calc.calculate(work)
        return calc.results["energy"], calc.results["forces"]

    host = md.velocity_verlet(ef, atoms.get_positions(), v0, m, steps=200, dt_fs=0.5)
    for use_graph in (False, True):
        dev = md.DeviceMD(calc.model, atoms.numbers, atoms.get_positions(), v0, m, dt_fs=0.5, use_graph=use_graph)
        out = dev.run(120)
        out = dev.run(80)          # a second call continues the same trajectory
        assert len(out["total"]) == 201
        assert np.abs(out["total"] - host["total"]).max() < 2e-4
        assert np.abs(out["potential"][:5] - host["potential"][:5]).max() < 5e-5
        x, v = dev.state()
        assert np.abs(x - host["positions"]).max() < 1e-4
    assert abs(out["drift_percent"]) < 0.14


def _nve_fixture():
    path = GOLDEN / "nve_reference.npz"
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def _device_run(calc, numbers, ref, key, steps):
    dev = md.DeviceMD(calc.model, numbers, ref[key + "_x0"], ref[key + "_v0"], ref[key + "_masses"], dt_fs=0.5)
    return dev.run(steps)


def test_nve_10ps_benzene_on_device_within_reference_bar(calc):
    """BASELINE north_star: NVE drift over a 10 ps velocity-Verlet run (20 000 x 0.5 fs, 300 K, COM
    translation + rotation removed: nve_harness.py:158-165, 329-331) within the reference's 0.14 %.
    Same initial state as the CPU reference trajectory of tests/golden/make_nve_reference.py: the
    end-point drifts are compared and the first 1 ps of the total-energy series must track it."""
    ref = _nve_fixture()
    key = "benzene_seed42"
    out = _device_run(calc, synthetic.benzene().numbers, ref, key, int(ref["steps"]))
    tot = out["total"]
    assert len(tot) == int(ref["steps"]) + 1
    assert abs(out["drift_percent"]) <= 0.14, out["drift_percent"]
    assert abs(float(ref[key + "_drift_percent"])) <= 0.14          # the CPU reference meets its own bar too
    assert np.abs(tot[:2001] - ref[key + "_total_first2001"]).max() < 2e-4
    band = 100.0 * float(np.max(np.abs(tot - tot[0]))) / abs(tot[0])
    assert band <= max(0.05, 2.0 * float(ref[key + "_band_percent"]))
    _write_nve_report("benzene", {"gpu_drift_percent": out["drift_percent"], "cpu_drift_percent": float(ref[key + "_drift_percent"]),
                                  "gpu_band_percent": band, "cpu_band_percent": float(ref[key + "_band_percent"])})


def _write_nve_report(name, payload):
    import json
    from conftest import ROOT
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"nve_10ps_{name}.json").write_text(json.dumps(payload, indent=1))


def test_nve_10ps_h2o_ensemble_inside_the_cpu_reference_band(calc):
    """H2O over 10 ps, 32 seeds, each started from the SAME (x0, v0) as a CPU run of the reference path
    (tests/golden/make_nve_reference.py).

    Finding (numbers in gpurun_out/nve_10ps_h2o.json): the reference model itself misses its 0.14 % bar on
    H2O -- most of the CPU reference's own end-point drifts are outside it -- because the total energy of
    this three-atom trajectory oscillates by 0.2 - 0.6 % of |E| under velocity Verlet at 0.5 fs and the
    metric samples that oscillation at one instant; two copies of one trajectory decorrelate within a
    fraction of a picosecond.  The on-device trajectories are therefore held to the reference's own
    DISTRIBUTION over the same 32 initial states: the start of every series tracks the CPU series, and
    the RMS and the 90th percentile of |drift| and the median oscillation band are no larger than the
    CPU ensemble's (+30 %, the sampling noise of a 32-member ensemble)."""
    ref = _nve_fixture()
    steps = int(ref["steps"])
    numbers = synthetic.water().numbers
    rows = []
    for seed in [int(s) for s in ref["h2o_seeds"]]:
        key = f"h2o_seed{seed}"
        out = _device_run(calc, numbers, ref, key, steps)
        tot = out["total"]
        band = 100.0 * float(np.max(np.abs(tot - tot[0]))) / abs(tot[0])
        diff = np.abs(tot[:2001] - ref[key + "_total_first2001"])
        rows.append({"seed": seed, "gpu_drift_percent": out["drift_percent"], "cpu_drift_percent": float(ref[key + "_drift_percent"]),
                     "gpu_band_percent": band, "cpu_band_percent": float(ref[key + "_band_percent"]),
                     "early_max_dE_eV": {"20_steps": float(diff[:21].max()), "100_steps": float(diff[:101].max()),
                                         "500_steps": float(diff[:501].max()), "2000_steps": float(diff.max())}})
    gpu = np.array([r["gpu_drift_percent"] for r in rows])
    cpu = np.array([r["cpu_drift_percent"] for r in rows])
    gband = np.array([r["gpu_band_percent"] for r in rows])
    cband = np.array([r["cpu_band_percent"] for r in rows])
    summary = {"seeds": len(rows),
               "rms_drift_percent": {"gpu": float(np.sqrt(np.mean(gpu ** 2))), "cpu": float(np.sqrt(np.mean(cpu ** 2)))},
               "p90_abs_drift_percent": {"gpu": float(np.percentile(np.abs(gpu), 90)), "cpu": float(np.percentile(np.abs(cpu), 90))},
               "mean_drift_percent": {"gpu": float(gpu.mean()), "cpu": float(cpu.mean())},
               "median_band_percent": {"gpu": float(np.median(gband)), "cpu": float(np.median(cband))},
               "within_0.14_percent": {"gpu": int(np.sum(np.abs(gpu) <= 0.14)), "cpu": int(np.sum(np.abs(cpu) <= 0.14))}}
    _write_nve_report("h2o", {"summary": summary, "rows": rows})
    for r in rows:
        assert r["early_max_dE_eV"]["20_steps"] < 1e-4, r       # same forces, same integrator: the start coincides
    assert summary["within_0.14_percent"]["cpu"] < len(rows) // 2    # the finding: the reference misses its own bar here
    assert summary["rms_drift_percent"]["gpu"] <= 1.3 * summary["rms_drift_percent"]["cpu"], summary
    assert summary["p90_abs_drift_percent"]["gpu"] <= 1.3 * summary["p90_abs_drift_percent"]["cpu"], summary
    assert summary["median_band_percent"]["gpu"] <= 1.3 * summary["median_band_percent"]["cpu"], summary


def test_device_md_freezes_and_resumes_on_edge_overflow(calc):
    """A force evaluation that overflows the edge workspace mid-trajectory must not corrupt the state:
    the guarded kick / drift kernels freeze at that step, DeviceMD.run grows the workspace, completes
    the step and carries on -- same series as a run with ample capacity, bit for bit.  Two benzene
    molecules start beyond each other's cutoff and fly together, so the edge count must grow."""
    from mlff_distiller_b200.student_model import StudentForceField
    a, b = synthetic.benzene(), synthetic.benzene()
    numbers = np.concatenate([a.numbers, b.numbers])
    pos = np.concatenate([a.get_positions(), b.get_positions() + np.array([10.5, 0.0, 0.0])])
    m = md.ATOMIC_MASSES[numbers]
    v0 = np.zeros_like(pos)
    v0[:12, 0], v0[12:, 0] = 0.25, -0.25           # Angstrom per ASE time unit: ~0.025 A closer per step
    steps = 120
    ample = md.DeviceMD(calc.model, numbers, pos, v0, m, dt_fs=0.5, edge_reserve=2000)
    e_start = int(ample.eng.status().num_edges)
    ref = ample.run(steps)
    assert ample.interruptions == 0 and int(ample.eng.status().num_edges) > e_start + 2
    tight_model = StudentForceField.load(GOLDEN / "weights_original.npz", device="cuda")
    tight_model.engine().reserve(len(numbers), e_start + 2, 1)   # exact: the first new pair overflows it
    dev = md.DeviceMD(tight_model, numbers, pos, v0, m, dt_fs=0.5, edge_reserve=1)
    assert dev.eng.cap_edges == e_start + 2
    out = dev.run(steps)
    assert dev.interruptions >= 1
    assert len(out["total"]) == steps + 1
    assert np.array_equal(out["total"], ref["total"])


def test_fp16_range_guard_falls_back_to_the_fp32_kernels():
    """ADVICE / VERDICT weak #5: the two-term FP16 split has a finite range.  (a) activations: with
    the embedding scaled x 1e4 the update block's operands leave the FP16 range; the producers raise
    tc_saturated and the call is repeated on the FP32 FFMA kernels -- the result equals the fp32 model's.
    (b) weights: a matrix with max|w| >= 253 cannot be pre-scaled; such a model never uses the tensor cores."""
    from conftest import load_weights
    from mlff_distiller_b200.checkpoint import infer_config
    from mlff_distiller_b200.student_model import StudentForceField
    state, cfg = load_weights("original")
    structs = synthetic.druglike_batch(48, first=50)   # 2 400 atoms: above MLFFD_SMALL_ROWS, so the tcgen05 update block runs
    z, pos, off = synthetic.concatenate(structs)
    z_d = torch.from_numpy(z.astype(np.int32)).cuda()
    p_d = torch.from_numpy(pos.astype(np.float32)).cuda()
    o_d = torch.from_numpy(off.astype(np.int32)).cuda()

    def run(st, precision):
        model = StudentForceField.from_state(st, infer_config(st, cfg), "cuda:0", precision=precision)
        e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(structs))
        return model, e.cpu().numpy(), f.cpu().numpy()

    hot = dict(state)
    hot["embedding.weight"] = state["embedding.weight"] * 1.0e4
    m_tc, e_tc, f_tc = run(hot, "tc")
    m_32, e_32, f_32 = run(hot, "fp32")
    assert m_tc.engine().saturation_reruns == 1
    assert np.isfinite(e_tc).all() and np.isfinite(f_tc).all()
    assert np.array_equal(e_tc, e_32) and np.array_equal(f_tc, f_32)
    m_ok, e_ok, _ = run(state, "tc")
    assert m_ok.engine().saturation_reruns == 0            # trained weights stay far inside the range
    big = dict(state)
    key = "interactions.2.update.update_mlp.2.weight"
    big[key] = state[key] * np.float32(300.0 / np.abs(state[key]).max())      # max|w| = 300 >= 253
    m_w, e_w, f_w = run(big, "tc")
    m_w32, e_w32, f_w32 = run(big, "fp32")
    assert m_w.engine().saturation_reruns == 0
    assert np.isfinite(e_w).all() and np.isfinite(f_w).all()
    assert np.array_equal(e_w, e_w32) and np.array_equal(f_w, f_w32)


def test_device_ids_spread_one_batch_over_the_gpus_of_the_process():
    """SURVEY section 8e / VERDICT missing #4: one process, several GPUs behind the calculator API --
    calculate_batch / evaluate_arrays split one structure list over device_ids (no torchrun, no collective)
    and return the same results, in input order, as a single device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in the process")
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    ids = list(range(min(torch.cuda.device_count(), 4)))
    multi = StudentForceFieldCalculator(GOLDEN / "weights_original.npz", device_ids=ids)
    single = StudentForceFieldCalculator(GOLDEN / "weights_original.npz", device="cuda:0")
    structs = synthetic.druglike_batch(37, first=4000, ragged=True)
    z, pos, off = synthetic.concatenate(structs)
    e_m, f_m = multi.evaluate_arrays(z, pos, np.diff(off))
    e_s, f_s = single.evaluate_arrays(z, pos, np.diff(off))
    # the kernels a call takes depend on its size (small-system path, tile shapes), so a shard and the whole
    # list agree to rounding, not bit for bit
    assert np.abs(e_m - e_s).max() / 80 <= 1e-6 and np.abs(f_m - f_s).max() <= 2e-5
    res = multi.calculate_batch(structs)
    assert len(res) == len(structs)
    assert all(r["forces"].shape == (len(s), 3) for r, s in zip(res, structs))
    assert np.allclose([r["energy"] for r in res], e_s)
    assert torch.cuda.current_device() == 0      # the library leaves the caller's current device alone


def test_device_md_batch_of_independent_trajectories(calc):
    structs = [synthetic.water(), synthetic.benzene(), synthetic.druglike(5, 30)]
    z, pos, off = synthetic.concatenate(structs)
    m = md.ATOMIC_MASSES[z]
    v0 = md.maxwell_boltzmann(m, 300.0, np.random.default_rng(1))
    dev = md.DeviceMD(calc.model, z, pos, v0, m, dt_fs=0.5, offsets=off)
    out = dev.run(100)
    assert len(out["total"]) == 101 and np.isfinite(out["total"]).all()
    assert abs(out["drift_percent"]) < 0.5


def test_batched_interface_micro_batches_large_requests(calc):
    """A request larger than the per-call atom budget is split into micro-batches (C5 sweeps);
    results must not depend on where the cuts fall."""
    structs = synthetic.druglike_batch(40, first=2000, ragged=True)
    z, pos, off = synthetic.concatenate(structs)
    counts = np.diff(off)
    e_ref, f_ref = calc.evaluate_arrays(z, pos, counts)
    old = calc.max_atoms_per_call
    try:
        calc.max_atoms_per_call = 300
        e, f = calc.evaluate_arrays(z, pos, counts)
    finally:
        calc.max_atoms_per_call = old
    assert e.shape == e_ref.shape and f.shape == f_ref.shape
    assert np.abs(e - e_ref).max() < 2e-4 and np.abs(f - f_ref).max() < 2e-5


def _ragged_batches(num, per, first):
    out = []
    for b in range(num):
        structs = synthetic.druglike_batch(per, first=first + b * per, ragged=True)
        z, pos, off = synthetic.concatenate(structs)
        out.append((z, pos, np.diff(off)))
    return out


def test_evaluate_arrays_fills_caller_provided_outputs():
    """``out=`` (a rank's slice of a shared-memory result in the multi-GPU sweep): same values, written in
    place, for the single-batch and the micro-batched path."""
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    calc = StudentForceFieldCalculator(GOLDEN / "weights_ultra_tiny.npz", device="cuda")
    structs = synthetic.druglike_batch(40, first=900, ragged=True)
    z, pos, off = synthetic.concatenate(structs)
    counts = np.diff(off)
    e_ref, f_ref = calc.evaluate_arrays(z, pos, counts)
    for budget in (calc.max_atoms_per_call, 400):
        calc.max_atoms_per_call = budget
        e_out = np.full(len(counts), np.nan, dtype=np.float32)
        f_out = np.full((len(z), 3), np.nan, dtype=np.float32)
        e, f = calc.evaluate_arrays(z, pos, counts, out=(e_out, f_out))
        assert e is e_out and f is f_out
        assert np.allclose(e_out, e_ref, rtol=0, atol=2e-5) and np.allclose(f_out, f_ref, rtol=0, atol=2e-5)
    with pytest.raises(ValueError):
        calc.evaluate_arrays(z, pos, counts, out=(np.zeros(3, np.float32), f_out))


def test_evaluate_stream_matches_blocking_interface(calc):
    """The pipelined sweep interface (two batches in flight, copies under the kernels) returns
    exactly what the blocking call returns, batch by batch and in order."""
    batches = _ragged_batches(5, 24, first=3000) + _ragged_batches(2, 3, first=3500)
    ref = [calc.evaluate_arrays(*b) for b in batches]
    n0 = calc.n_calls
    got = list(calc.evaluate_stream(iter(batches)))
    assert calc.n_calls == n0 + len(batches)
    assert len(got) == len(ref)
    for (e, f), (e_ref, f_ref), b in zip(got, ref, batches):
        assert e.dtype == np.float32 and f.dtype == np.float32
        assert e.shape == (len(b[2]),) and f.shape == (len(b[0]), 3)
        assert np.array_equal(e, e_ref) and np.array_equal(f, f_ref)   # deterministic kernels
    assert list(calc.evaluate_stream(iter([]))) == []
    with pytest.raises(ValueError):
        list(calc.evaluate_stream(iter([(np.array([1, 200]), np.zeros((2, 3)), np.array([2]))])))


def test_evaluate_stream_recovers_from_edge_overflow():
    """A batch whose edges overflow the workspace guess is detected from the status words that
    travel with its results and re-run; the batches around it are unaffected."""
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    fresh = StudentForceFieldCalculator(GOLDEN / "weights_original.npz", device="cuda")
    fresh.model._edges_per_atom = 1   # first workspace guess far too small -> overflow in flight
    batches = _ragged_batches(4, 16, first=4000)
    got = list(fresh.evaluate_stream(iter(batches)))
    good = StudentForceFieldCalculator(GOLDEN / "weights_original.npz", device="cuda")
    for (e, f), b in zip(got, batches):
        e_ref, f_ref = good.evaluate_arrays(*b)
        assert np.array_equal(e, e_ref) and np.array_equal(f, f_ref)


def test_use_jit_reads_the_torchscript_archive_and_wrapper_differentiates(calc, tmp_path):
    """`use_jit=True`: the weights come from the TorchScript archive; the (Z, R) -> E wrapper
    gives forces through autograd exactly as inference/ase_calculator.py:319-335 computes them."""
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    from mlff_distiller_b200.student_model import EnergyOnlyWrapper, StudentForceField
    from test_checkpoint import _export_like_reference
    cpu_model = StudentForceField.load(GOLDEN / "weights_original.npz")
    _export_like_reference(cpu_model, tmp_path / "student_jit.pt")
    jit_calc = StudentForceFieldCalculator(tmp_path / "unused.pt", device="cuda", use_jit=True,
                                           jit_path=tmp_path / "student_jit.pt")
    a, b = synthetic.benzene(), synthetic.benzene()
    a.calc, b.calc = jit_calc, calc
    assert a.get_potential_energy() == b.get_potential_energy()
    assert np.array_equal(a.get_forces(), b.get_forces())
    wrapper = EnergyOnlyWrapper(calc.model)
    z = torch.from_numpy(a.get_atomic_numbers()).to("cuda")
    pos = torch.from_numpy(a.get_positions().astype(np.float32)).to("cuda").requires_grad_(True)
    energy = wrapper(z, pos)
    forces = -torch.autograd.grad(energy, pos)[0]
    assert abs(float(energy) - b.get_potential_energy()) < 1e-5
    assert np.abs(forces.cpu().numpy() - b.get_forces()).max() < 1e-6


def test_single_structure_graph_replay_matches_eager(calc):
    """From the second call on a single-structure step is one CUDA-graph replay; it must return
    exactly what the eager launch sequence returns, follow changing positions, switch systems,
    and survive a workspace regrowth caused by another (larger) request in between."""
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    eager = StudentForceFieldCalculator(GOLDEN / "weights_original.npz", device="cuda", use_graph=False)
    rng = np.random.default_rng(3)
    mol = synthetic.druglike_batch(1, first=11)[0]
    base = mol.get_positions().copy()
    replays = 0
    for step in range(16):
        mol.set_positions(base + rng.normal(0, 0.02, base.shape))
        a, b = mol.copy(), mol.copy()
        a.calc, b.calc = calc, eager
        assert a.get_potential_energy() == b.get_potential_energy(), step
        assert np.array_equal(a.get_forces(), b.get_forces()), step
        replays += bool(calc._numbers_cache["graphs"])
        if step == 5:    # another system in between: new buffers, the capture count starts again
            w = synthetic.water()
            w.calc = calc
            w.get_forces()
        if step == 11:   # a big batch in between: regrown workspace, the captured pointers are stale
            calc.calculate_batch(synthetic.druglike_batch(64, first=500))
    assert replays >= 6, "the replay path was hardly taken"


def test_calculate_with_stage_profiling_enabled(calc):
    """Per-kernel profiling events cannot be recorded inside a graph capture: while profiling is on
    the single-structure path stays eager (and still returns the same numbers)."""
    atoms = synthetic.benzene()
    atoms.calc = calc
    ref = atoms.get_forces().copy()
    eng = calc.model.engine()
    eng.profile_enable(True)
    try:
        for k in range(3):
            b = synthetic.benzene()
            b.set_positions(b.get_positions() + 1e-3 * k)   # translation: same forces, new call
            b.calc = calc
            assert np.abs(b.get_forces() - ref).max() < 2e-5
        prof = eng.profile_read()
    finally:
        eng.profile_enable(False)
    assert prof["launches"] > 0 and prof["stages"]["message_fwd"]["ms"] > 0.0
