"""Host-side logic that needs no GPU: sharding, MD driver, synthetic workloads, validation."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from mlff_distiller_b200 import md, sharding, synthetic


def test_partition_covers_everything_once():
    rng = np.random.default_rng(0)
    counts = rng.integers(20, 81, size=1000)
    for shards in (1, 2, 4, 8, 3):
        parts = sharding.partition_by_atoms(counts, shards)
        assert parts[0][0] == 0 and parts[-1][1] == len(counts)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(shards - 1))
        loads = [counts[a:b].sum() for a, b in parts]
        assert max(loads) - min(loads) <= 2 * counts.max()
    assert sharding.partition_by_atoms([5, 5], 4)[-1][1] == 2
    chunks = sharding.chunk_by_budget(counts, max_atoms=4096, max_structs=64)
    assert chunks[0][0] == 0 and chunks[-1][1] == len(counts)
    assert all(counts[a:b].sum() <= 4096 and b - a <= 64 for a, b in chunks)


def test_partition_and_chunking_properties():
    """Property tests (hypothesis): for ANY list of structure sizes and ANY shard count the shards are
    contiguous, disjoint and cover the list (empty shards allowed when there are more ranks than
    structures), no shard exceeds the ideal load by more than the largest structure, and micro-batches
    respect both budgets except for a single structure that is larger than the atom budget by itself."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.integers(1, 400), min_size=0, max_size=200), st.integers(1, 16))
    def check_partition(counts, shards):
        parts = sharding.partition_by_atoms(counts, shards)
        assert len(parts) == shards and parts[0][0] == 0 and parts[-1][1] == len(counts)
        assert all(a <= b for a, b in parts)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(shards - 1))
        if counts:
            loads = [sum(counts[a:b]) for a, b in parts]
            assert sum(loads) == sum(counts)
            assert max(loads) <= sum(counts) / shards + max(counts)
        for r in range(shards):
            assert sharding.shard_slice(counts, r, shards) == parts[r]

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.integers(1, 400), min_size=0, max_size=200), st.integers(1, 2000), st.integers(1, 50))
    def check_chunks(counts, max_atoms, max_structs):
        chunks = sharding.chunk_by_budget(counts, max_atoms, max_structs)
        if not counts:
            assert chunks == []
            return
        assert chunks[0][0] == 0 and chunks[-1][1] == len(counts)
        assert all(chunks[i][1] == chunks[i + 1][0] for i in range(len(chunks) - 1))
        for a, b in chunks:
            assert 1 <= b - a <= max_structs
            assert sum(counts[a:b]) <= max_atoms or b - a == 1

    check_partition()
    check_chunks()


def test_synthetic_workloads_are_seeded_and_sane():
    a = synthetic.druglike_batch(3)
    b = synthetic.druglike_batch(3)
    for x, y in zip(a, b):
        assert np.array_equal(x.positions, y.positions) and np.array_equal(x.numbers, y.numbers)
    z, pos, off = synthetic.concatenate(a)
    assert list(off) == [0, 50, 100, 150] and synthetic.min_pair_distance(pos, off) >= 0.95
    r = synthetic.druglike_batch(20, ragged=True)
    assert {len(s) for s in r} != {50} and all(20 <= len(s) <= 80 for s in r)
    chain = synthetic.alkane_chain(100)
    assert len(chain) == 300 and list(chain.numbers[:3]) == [6, 1, 1]
    box = synthetic.water_box(64, seed=1)
    assert len(box) == 192 and box.pbc.all() and abs(box.cell[0, 0] - (64 / 0.0334) ** (1 / 3)) < 1e-9
    w = synthetic.water()
    assert np.allclose(w.positions[1], [0, 0.763239, -0.477047])


def test_velocity_verlet_conserves_energy_on_a_harmonic_toy():
    k = 5.0

    def force_fn(x):
        return 0.5 * k * float(np.sum(x ** 2)), -k * x

    rng = np.random.default_rng(3)
    masses = np.array([12.011, 1.008, 1.008])
    x0 = rng.normal(size=(3, 3)) * 0.1
    v0 = md.maxwell_boltzmann(masses, 300.0, rng)
    assert np.abs((masses[:, None] * v0).sum(0)).max() < 1e-12
    out = md.velocity_verlet(force_fn, x0, v0, masses, steps=2000, dt_fs=0.1)
    assert abs(out["drift_percent"]) < 0.02
    assert len(out["total"]) == 2001
    assert abs(md.FS - 0.09822694788) < 1e-9
    assert abs(md.ns_per_day(1000.0, 0.5) - 43.2) < 1e-9


def test_sharded_gather_world_size_2_gloo(tmp_path):
    """N>1 host logic on CPU: two gloo ranks take their shard, produce per-structure results with a
    stand-in evaluator and gather them back in input order."""
    script = tmp_path / "worker.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {str(ROOT)!r})
import numpy as np
import torch.distributed as dist
from mlff_distiller_b200 import sharding, synthetic
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
structs = synthetic.druglike_batch(11, ragged=True)
counts = [len(s) for s in structs]
a, b = sharding.shard_slice(counts, rank, world)
z, pos, off = synthetic.concatenate(structs[a:b])
e_local = np.array([pos[off[i]:off[i+1]].sum() for i in range(b - a)])   # stand-in "energy"
f_local = pos * 2.0                                                        # stand-in "forces"
e, f = sharding.gather_in_order(e_local, f_local, counts)
zz, pp, oo = synthetic.concatenate(structs)
assert e.dtype == np.float32 and f.dtype == np.float32
assert np.allclose(e, [pp[oo[i]:oo[i+1]].sum() for i in range(len(structs))], rtol=1e-6)
assert np.allclose(f, pp * 2.0, rtol=1e-6)
e2, f2 = sharding.gather_in_order(e_local, np.zeros((0, 3)), counts)      # energies only (sweeps)
assert np.array_equal(e2, e) and f2.shape == (0, 3)
e3, f3 = sharding.gather_in_order(e_local, f_local, counts, root=0)        # point-to-point into rank 0's output
if rank == 0:
    assert np.array_equal(e3, e) and np.array_equal(f3, f)
else:
    assert e3 is None and f3 is None
shared = sharding.SharedResults(counts, root=0)                            # one node: shared-memory gather of host results
for rep in range(2):
    shared.write(e_local + rep, f_local)
    e4, f4 = shared.collect()
    if rank == 0:
        assert np.allclose(e4, e + rep, rtol=1e-6) and np.array_equal(f4, f)
    else:
        assert e4 is None and f4 is None
    dist.barrier()
shared.close()
if rank == 0:
    print('GATHER_OK', a, b)
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                           "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
                           str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "GATHER_OK" in proc.stdout
    # only the root owns (and unlinks) the shared-memory segment: no second unlink by an attached rank's tracker
    assert "resource_tracker" not in proc.stderr, proc.stderr[-2000:]


def test_shared_results_single_process():
    """Without a process group SharedResults degenerates to one rank owning the whole list."""
    from mlff_distiller_b200 import sharding
    counts = [3, 5, 2]
    with sharding.SharedResults(counts) as shared:
        assert (shared.a, shared.b, shared.atom0, shared.atom1) == (0, 3, 0, 10)
        shared.write(np.arange(3), np.arange(30).reshape(10, 3))
        e, f = shared.collect()
        assert e.dtype == np.float32 and np.array_equal(e, [0, 1, 2]) and np.array_equal(f, np.arange(30).reshape(10, 3))
        del e, f
    shared.close()   # idempotent


def test_calculator_validation_messages_without_gpu():
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator

    class Fake:
        max_z = 100
        cutoff = 5.0

    calc = object.__new__(StudentForceFieldCalculator)
    calc.model = Fake()
    calc.pbc_mode = "minimum_image"
    ok = (np.zeros((2, 3)), np.array([1, 8]), np.eye(3) * 20, np.array([False] * 3))
    calc._validate_inputs(*ok)
    with pytest.raises(ValueError, match="empty structure"):
        calc._validate_inputs(np.zeros((0, 3)), np.array([], dtype=int), np.eye(3), np.array([False] * 3))
    with pytest.raises(ValueError, match="Invalid atomic numbers"):
        calc._validate_inputs(np.zeros((1, 3)), np.array([0]), np.eye(3), np.array([False] * 3))
    with pytest.raises(ValueError, match="Invalid atomic numbers"):
        calc._validate_inputs(np.zeros((1, 3)), np.array([110]), np.eye(3), np.array([False] * 3))
    with pytest.raises(ValueError, match="NaN"):
        calc._validate_inputs(np.array([[np.nan, 0, 0]]), np.array([1]), np.eye(3), np.array([False] * 3))
    with pytest.raises(ValueError, match="2\\*cutoff"):
        calc._validate_inputs(np.zeros((1, 3)), np.array([29]), np.eye(3) * 3.58, np.array([True] * 3))
    with pytest.raises(FileNotFoundError):
        StudentForceFieldCalculator("/nonexistent/best_model.pt", device="cuda")


def test_use_jit_flag_errors_like_the_reference(tmp_path):
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    with pytest.raises(ValueError, match="jit_path not provided"):
        StudentForceFieldCalculator(tmp_path / "x.pt", device="cuda", use_jit=True)
    with pytest.raises(FileNotFoundError, match="TorchScript model not found"):
        StudentForceFieldCalculator(tmp_path / "x.pt", device="cuda", use_jit=True, jit_path=tmp_path / "missing_jit.pt")


def test_validate_arrays_messages_without_gpu():
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator

    class Fake:
        max_z = 100

    calc = object.__new__(StudentForceFieldCalculator)
    calc.model = Fake()
    z, pos = np.array([1, 8, 1, 6]), np.zeros((4, 3))
    calc._validate_arrays(z, pos, np.array([3, 1]))
    with pytest.raises(ValueError, match="empty structure"):
        calc._validate_arrays(z, pos, np.array([4, 0]))
    with pytest.raises(ValueError, match="disagree"):
        calc._validate_arrays(z, pos, np.array([2, 1]))
    with pytest.raises(ValueError, match="Invalid atomic numbers"):
        calc._validate_arrays(np.array([1, 8, 1, 101]), pos, np.array([4]))
    with pytest.raises(ValueError, match="NaN"):
        calc._validate_arrays(z, np.full((4, 3), np.inf), np.array([4]))


def _reference_arm(env_extra):
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env={**os.environ, **env_extra})
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver parses stdout of bench.py: one JSON line, nothing else (library banners go to stderr).
    With neither the reference tree nor the packed archive the arm is the restatement (kind "port")."""
    d = _reference_arm({"MLFFD_BENCH_FORCE_PORT": "1"})
    assert d["impl"] == "reference" and d["metric"] == "structures_per_second_energy_forces"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("C2") and d["higher_is_better"] is True


def test_reference_archive_is_the_reference_and_the_arm_runs_it():
    """oracle/build_ref.py packs the reference's two files, unmodified, into oracle/_ref/reference_path.zip (what
    travels to the GPU box); imported from there with zipimport the model gives the outputs of the restatement
    bit for bit, and bench.py --impl reference then reports kind "reference"."""
    import hashlib
    import json
    import zipfile
    import torch
    from conftest import load_weights
    from mlff_distiller_b200 import synthetic
    from oracle import build_ref, reference_loader
    import oracle.painn_oracle as po
    path = build_ref.build()
    if path is None:
        pytest.skip("neither the reference tree nor a prebuilt archive is present")
    with zipfile.ZipFile(path) as z:
        manifest = json.loads(z.read("MANIFEST.json"))["sha256"]
        for member, digest in manifest.items():
            assert hashlib.sha256(z.read(member[len("src/"):])).hexdigest() == digest
            if reference_loader.available():
                assert (reference_loader.REFERENCE_ROOT / member).read_bytes() == z.read(member[len("src/"):])
        assert z.read("mlff_distiller/__init__.py") == b"" and z.read("mlff_distiller/models/__init__.py") == b""
    d = _reference_arm({"MLFFD_REFERENCE_SOURCE": "archive"})
    assert d["cpu_baseline"]["kind"] == "reference" and "archive" in d["cpu_baseline"]["sample"]
    assert d["gpu_launches"] == 0 and d["value"] > 0

    # same outputs as the restatement, in a subprocess so that this process keeps whatever it imported before
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import sys, json, numpy as np, torch; sys.path.insert(0, %r)\n"
        "from types import SimpleNamespace\n"
        "from oracle import reference_loader as rl, painn_oracle as po\n"
        "from mlff_distiller_b200 import synthetic\n"
        "z = np.load(%r); state = {k: z[k] for k in z.files if not k.startswith('__')}; cfg = json.loads(str(z['__config__']))\n"
        "torch.set_num_threads(1)   # one thread: the scatter-adds of both implementations accumulate in the same order\n"
        "assert rl.source() == 'archive'\n"
        "model = rl.build_reference_model(state, SimpleNamespace(**cfg))\n"
        "assert '.zip' in sys.modules['mlff_distiller.models.student_model'].__file__\n"
        "zz, pos, off = synthetic.concatenate(synthetic.druglike_batch(3, first=99, ragged=True))\n"
        "p = torch.from_numpy(pos.astype(np.float32)).requires_grad_(True); b = po.batch_from_offsets(off)\n"
        "e = model(atomic_numbers=torch.from_numpy(zz), positions=p, cell=None, pbc=None, batch=b)\n"
        "f = -torch.autograd.grad(e, p, grad_outputs=torch.ones_like(e))[0]\n"
        "e2, f2 = po.energy_and_forces(po.to_torch_weights(state), torch.from_numpy(zz), torch.from_numpy(pos.astype(np.float32)), cfg['cutoff'], b)\n"
        "assert torch.equal(e.detach(), e2) and torch.equal(f, f2), (float((e.detach()-e2).abs().max()), float((f-f2).abs().max()))\n"
        "print('identical')\n") % (str(ROOT), str(ROOT / "tests" / "golden" / "weights_ultra_tiny.npz"))
    import os
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                          env={**os.environ, "MLFFD_REFERENCE_SOURCE": "archive"})
    assert proc.returncode == 0 and "identical" in proc.stdout, proc.stderr[-2000:]


def test_calculate_batch_marshalling_without_gpu():
    """List of structure objects -> stacked arrays -> list of result dicts (ase_calculator.py:647-706, :765-817),
    with a stand-in for the device evaluation: atoms-like objects with ASE's `numbers` / `positions`
    attributes and ones that only offer the getters give the same stacked inputs; results are cut back per
    structure in input order with the reference's keys."""
    from mlff_distiller_b200 import synthetic
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator

    class GettersOnly:
        def __init__(self, s):
            self._s = s

        def __len__(self):
            return len(self._s)

        def get_atomic_numbers(self):
            return self._s.get_atomic_numbers()

        def get_positions(self):
            return self._s.get_positions()

        def get_pbc(self):
            return self._s.get_pbc()

        def get_cell(self):
            return self._s.get_cell()

    calc = object.__new__(StudentForceFieldCalculator)
    calc.pbc_mode, calc.enable_stress = "ignore", True
    seen = []

    def fake_evaluate(numbers, positions, counts, cells=None, pbcs=None, out=None):
        seen.append((numbers.copy(), positions.copy(), counts.copy(), cells, pbcs))
        offs = np.concatenate([[0], np.cumsum(counts)])
        e = np.array([positions[offs[i]:offs[i + 1]].sum() for i in range(len(counts))], dtype=np.float32)
        return e, (positions * 2.0).astype(np.float32)

    calc.evaluate_arrays = fake_evaluate
    structs = synthetic.druglike_batch(7, first=3, ragged=True)
    res_a = calc.calculate_batch(structs)
    res_g = calc.calculate_batch([GettersOnly(s) for s in structs])
    for a, b in zip(seen[0][:3], seen[1][:3]):
        assert np.array_equal(a, b)
    z, pos, off = synthetic.concatenate(structs)
    assert np.array_equal(seen[0][0], z) and np.array_equal(seen[0][1], pos) and seen[0][3] is None
    assert np.array_equal(seen[0][2], np.diff(off))
    for res in (res_a, res_g):
        assert len(res) == len(structs)
        for i, (r, s) in enumerate(zip(res, structs)):
            assert set(r) == {"energy", "forces"} and isinstance(r["energy"], float)
            assert r["energy"] == float(np.float32(s.positions.sum()))
            assert r["forces"].shape == (len(s), 3) and np.array_equal(r["forces"], (s.positions * 2.0).astype(np.float32))
    only_e = calc.calculate_batch(structs[:2], properties=["energy"])
    assert [set(r) for r in only_e] == [{"energy"}, {"energy"}]
    with_s = calc.calculate_batch(structs[:2], properties=["energy", "forces", "stress"])
    assert all(r["stress"] is None for r in with_s)      # "not supported in batch mode yet", like the reference
    assert calc.calculate_batch([]) == []
    with pytest.raises(ValueError, match="empty structure"):
        calc.calculate_batch([structs[0], synthetic.Structure([], np.zeros((0, 3)))])


def test_minimum_image_check_is_shared_by_every_entry_point_that_takes_a_cell():
    """A periodic cell height below 2 (r_c + skin) would silently lose images in the one-image kernels: the
    calculator (single and batched), StudentForceField._prepare (forward / analytical forces / stress) and
    md.DeviceMD all run the same host-side check before anything reaches the device."""
    import torch
    from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
    from mlff_distiller_b200.student_model import StudentForceField, check_minimum_image
    sheared = np.eye(3) * 10.6
    sheared[1, 0], sheared[2, 1] = 0.7, -0.5
    check_minimum_image(sheared, [True] * 3, 5.0)                                   # heights 10.58 .. 10.6
    check_minimum_image(torch.from_numpy(sheared)[None], torch.tensor([[True] * 3]), 5.0)
    check_minimum_image(np.diag([20.0, 3.0, 20.0]), [True, False, True], 5.0)       # the short axis is not periodic
    check_minimum_image(np.eye(3) * 3, [False] * 3, 5.0)
    with pytest.raises(ValueError, match=r"2\*cutoff.*axis 0 has 9\.900"):
        check_minimum_image(np.eye(3) * 9.9, [True, False, False], 5.0)
    with pytest.raises(ValueError, match="structure 1, axis 2"):
        check_minimum_image(np.stack([np.eye(3) * 20, np.eye(3) * 8]), np.array([[True] * 3, [False, False, True]]), 5.0)
    with pytest.raises(ValueError, match=r"skin=1\.0 needs"):
        check_minimum_image(np.eye(3) * 11, [True] * 3, 5.0, 1.0)
    # a sheared cell whose edge lengths pass but whose height does not
    thin = np.array([[10.5, 0, 0], [9.0, 5.4, 0], [0, 0, 12.0]])
    assert np.linalg.norm(thin[1]) > 10.0
    with pytest.raises(ValueError, match=r"axis 0 has 5\.40"):
        check_minimum_image(thin, [True] * 3, 5.0)

    model = StudentForceField(hidden_dim=32, num_interactions=1, num_rbf=4, cutoff=5.0, max_z=10, pbc_mode="minimum_image")
    with pytest.raises(ValueError, match=r"2\*cutoff"):
        model._prepare(torch.tensor([1, 8]), torch.zeros(2, 3), torch.eye(3) * 6, torch.tensor([True] * 3), None)
    ignore = StudentForceField(hidden_dim=32, num_interactions=1, num_rbf=4, cutoff=5.0, max_z=10)
    ignore._prepare(torch.tensor([1, 8]), torch.zeros(2, 3), torch.eye(3) * 6, torch.tensor([True] * 3), None)   # pbc_mode='ignore'

    class Fake:
        max_z, cutoff = 100, 5.0

    calc = object.__new__(StudentForceFieldCalculator)
    calc.model, calc.pbc_mode, calc.skin, calc._replicas = Fake(), "minimum_image", 0.0, None
    z, pos, counts = np.array([1, 8, 1, 8]), np.zeros((4, 3)), np.array([2, 2])
    with pytest.raises(ValueError, match="structure 1, axis 0"):
        calc._evaluate_arrays(z, pos, counts, np.stack([np.eye(3) * 20, np.eye(3) * 7]), np.ones((2, 3), dtype=bool), None)


def test_bench_sweep_record_glue_with_a_stand_in_calculator(monkeypatch):
    """bench.py's one-GPU C5 record (ragged list -> evaluate_arrays -> structures/s per variant) and the guard that
    keeps a failing extra record from costing the headline line, exercised without a GPU."""
    import importlib.util
    import types
    import torch
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import mlff_distiller_b200.ase_calculator as ac
    calls = []

    class FakeCalc:
        max_atoms_per_call = 1500

        def __init__(self, path, device, precision, filter_mode):
            assert path.exists()
            calls.append(path.name)

        def evaluate_arrays(self, numbers, positions, counts):
            assert numbers.dtype == np.int64 and positions.dtype == np.float64 and int(counts.sum()) == len(numbers)
            offs = np.concatenate([[0], np.cumsum(counts)])
            e = np.add.reduceat(positions.sum(1), offs[:-1]).astype(np.float32)
            return e, (positions * 0.5).astype(np.float32)

    monkeypatch.setattr(ac, "StudentForceFieldCalculator", FakeCalc)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    args = types.SimpleNamespace(precision="tc", filter_mode="spline")
    rec = bench.sweep_record(args, "cpu", distinct=40, repeat=3, passes=2)
    assert rec["structures"] == 120 and rec["distinct_structures"] == 40 and rec["atoms"] % 3 == 0
    assert calls == ["weights_original.npz", "weights_tiny.npz", "weights_ultra_tiny.npz"]
    for v in ("original", "tiny", "ultra_tiny"):
        assert rec[v]["structures_per_s"] > 0 and rec[v]["max_energy_spread_between_repeats_eV"] == 0.0

    def boom():
        raise RuntimeError("no device")

    assert bench.guarded(boom) == {"error": "RuntimeError: no device"}
    assert bench.guarded(lambda x: x + 1, 1) == 2
