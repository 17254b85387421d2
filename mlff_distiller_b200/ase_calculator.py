"""``StudentForceFieldCalculator`` on the B200 CUDA path.

Mirrors the reference calculator (src/mlff_distiller/inference/ase_calculator.py:61-886, paths
relative to /root/reference): constructor keywords (:105-120), ``calculate`` with ASE result
caching (:279-412), input validation and its exception types/messages (:414-459),
``calculate_batch`` (:590-817), ``reset`` / ``n_calls`` / ``avg_time`` / ``get_timing_stats``
(:819-874) and ``__repr__`` (:876-883).  When ASE is installed the class derives from
``ase.calculators.calculator.Calculator``; otherwise a small stand-in base provides the same
caching contract so MD drivers and tests run without ASE.

Everything numeric happens in libmlffd.so; the reference's optimisation switches
(``use_compile``, ``use_fp16``, ``use_jit``, ``use_torch_cluster``, ``use_analytical_forces``,
``batch_size``) are accepted for call-site compatibility and ignored with a log line.
"""
from __future__ import annotations

import logging
import time
import warnings
from pathlib import Path
from typing import Any, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .student_model import StudentForceField, check_minimum_image

logger = logging.getLogger(__name__)

try:  # pragma: no cover - ASE is optional
    from ase.calculators.calculator import Calculator as _AseCalculator, all_changes
    HAVE_ASE = True
except Exception:  # ImportError or a broken install
    HAVE_ASE = False
    all_changes = ["positions", "numbers", "cell", "pbc", "initial_charges", "initial_magmoms"]

    class _AseCalculator:  # minimal stand-in with ASE's caching contract
        implemented_properties: List[str] = []

        def __init__(self, **kwargs):
            self.atoms = None
            self.results: Dict[str, Any] = {}
            self.parameters = dict(kwargs)

        def reset(self):
            self.atoms = None
            self.results = {}

        def calculate(self, atoms=None, properties=("energy",), system_changes=all_changes):
            if atoms is not None:
                self.atoms = atoms.copy()

        def _changed(self, atoms) -> bool:
            a = self.atoms
            if a is None or len(a) != len(atoms):
                return True
            return not (np.array_equal(a.get_positions(), atoms.get_positions())
                        and np.array_equal(a.get_atomic_numbers(), atoms.get_atomic_numbers())
                        and np.array_equal(np.asarray(a.get_cell()), np.asarray(atoms.get_cell()))
                        and np.array_equal(a.get_pbc(), atoms.get_pbc()))

        def get_property(self, name, atoms=None):
            if atoms is None:
                atoms = self.atoms
            if name not in self.results or self._changed(atoms):
                self.calculate(atoms, [name], all_changes)
            return self.results[name]

        def get_potential_energy(self, atoms=None):
            return self.get_property("energy", atoms)

        def get_forces(self, atoms=None):
            return self.get_property("forces", atoms)

        def get_stress(self, atoms=None):
            return self.get_property("stress", atoms)


class StudentForceFieldCalculator(_AseCalculator):
    """ASE-style calculator backed by the hand-written CUDA energy+force path."""

    def __init__(self, checkpoint_path: Union[str, Path], device: str = "cuda",
                 dtype: torch.dtype = torch.float32, enable_stress: bool = False,
                 batch_size: Optional[int] = None, enable_timing: bool = False,
                 use_compile: bool = False, use_fp16: bool = False, use_jit: bool = False,
                 jit_path: Optional[Union[str, Path]] = None, use_torch_cluster: bool = True,
                 use_analytical_forces: bool = False, *, precision: str = "tc",
                 pbc_mode: str = "ignore", use_graph: bool = True, filter_mode: str = "spline",
                 device_ids: Optional[Sequence[int]] = None, skin: float = 0.0, **kwargs):
        super().__init__(**kwargs)
        self.checkpoint_path = Path(checkpoint_path)
        # device_ids=[0, 1, ...]: ONE process drives several GPUs: this instance owns the
        # first device, one replica (own context, streams and pinned staging) each further one; batched calls
        # split their structure list by atom count over the devices, no collective, results in input order
        self.device_ids = [int(i) for i in device_ids] if device_ids else None
        if self.device_ids:
            device = f"cuda:{self.device_ids[0]}"
        self.device = torch.device(device)
        self.dtype = dtype
        self.enable_stress = enable_stress
        self.batch_size = batch_size
        self.enable_timing = enable_timing
        self.use_compile = use_compile
        self.use_fp16 = use_fp16
        self.use_jit = use_jit
        self.jit_path = Path(jit_path) if jit_path else None
        self.use_torch_cluster = use_torch_cluster
        self.use_analytical_forces = use_analytical_forces
        self.precision = precision
        self.filter_mode = filter_mode
        self.skin = float(skin)   # Verlet-skin neighbour list for trajectories of one system (0 = off, the reference's behaviour)
        self.pbc_mode = pbc_mode
        self.use_graph = use_graph   # replay single-structure steps as one CUDA graph once a system keeps coming back
        self.graph_after_calls = 3   # eager calls for a system before its step is captured
        self.max_atoms_per_call = 262144   # micro-batch bound of the batched interface (~30 GB workspace)
        self.implemented_properties = ["energy", "forces"]
        if self.enable_stress:
            self.implemented_properties.append("stress")
        if dtype != torch.float32:
            raise ValueError("the CUDA path computes in float32; use the reference for other dtypes")
        ignored = [n for n, v in (("use_compile", use_compile), ("use_fp16", use_fp16),
                                  ("batch_size", batch_size)) if v]
        if ignored:
            logger.info("StudentForceFieldCalculator: ignoring %s (superseded by the CUDA path)", ignored)
        self.model = self._load_model()
        self._n_calls = 0
        self._total_time = 0.0
        self._call_times: List[float] = []
        self._numbers_cache = None  # per-system device / pinned buffers and captured graphs of the single path
        self._replicas = None
        if self.device_ids and len(self.device_ids) > 1:
            from concurrent.futures import ThreadPoolExecutor
            self._replicas = [self] + [
                StudentForceFieldCalculator(checkpoint_path, device=f"cuda:{i}", dtype=dtype, enable_stress=enable_stress,
                                            enable_timing=False, use_jit=use_jit, jit_path=jit_path, precision=precision,
                                            pbc_mode=pbc_mode, use_graph=use_graph, filter_mode=filter_mode)
                for i in self.device_ids[1:]]
            self._pool = ThreadPoolExecutor(max_workers=len(self._replicas))
        logger.info("Initialized StudentForceFieldCalculator: device=%s, precision=%s, pbc_mode=%s",
                    self.device, precision, pbc_mode)

    # ---- model ----------------------------------------------------------------------------
    def _load_model(self) -> StudentForceField:
        if self.use_jit:
            # inference/ase_calculator.py:216-236: the model comes from the TorchScript archive.
            # Its weights are read out of the archive and evaluated by the CUDA path (the archive's
            # own graph is the reference's eager ops; it is not executed).
            if not self.jit_path:
                raise ValueError("use_jit=True but jit_path not provided")
            if not self.jit_path.exists():
                raise FileNotFoundError(
                    f"TorchScript model not found: {self.jit_path}\n"
                    f"Please export model to TorchScript first using scripts/export_to_torchscript.py")
            source = self.jit_path
        else:
            source = self.checkpoint_path
        if not source.exists():
            raise FileNotFoundError(
                f"Checkpoint not found: {source}\n"
                f"Please ensure the model has been trained and checkpoint saved.")
        try:
            model = StudentForceField.load(source, device=str(self.device),
                                           precision=self.precision, pbc_mode=self.pbc_mode,
                                           filter_mode=self.filter_mode, skin=self.skin)
            model.eval()
            model.engine()  # fail loudly now if the CUDA library / device is missing
            return model
        except FileNotFoundError:
            raise
        except Exception as e:
            raise RuntimeError(f"Failed to load model from {source}: {e}") from e

    # ---- single structure -----------------------------------------------------------------
    def calculate(self, atoms=None, properties: Sequence[str] = ("energy", "forces"),
                  system_changes: Sequence[str] = all_changes):
        _AseCalculator.calculate(self, atoms, properties, system_changes)
        start = time.perf_counter() if self.enable_timing else 0.0
        try:
            positions = atoms.get_positions()
            numbers = atoms.get_atomic_numbers()
            cell = np.asarray(atoms.get_cell(), dtype=np.float64)
            pbc = np.asarray(atoms.get_pbc(), dtype=bool)
            self._validate_inputs(positions, numbers, cell, pbc)
            want_stress = "stress" in properties and self.enable_stress
            volume = abs(float(np.linalg.det(cell))) if want_stress else 0.0   # an MD loop asks for energy + forces only
            # Stress semantics.  The reference returns zeros without a cell or without a periodic axis
            # (ase_calculator.py:548-550) and, because its model never reads `cell`, zeros for periodic
            # input as well (the autograd call at :566 fails and :586-588 falls back).  pbc_mode='ignore' is
            # the mode that reproduces the reference, so it returns those zeros; with
            # pbc_mode='minimum_image' the model does see the cell and the stress is the real one:
            # (1/V) dE/d(strain) from the edge adjoints of the force evaluation.
            real_stress = (want_stress and self.pbc_mode == "minimum_image" and bool(pbc.any())
                           and volume > 1e-12)
            energy, forces, virial = self._evaluate_single(positions, numbers, cell, pbc, want_virial=real_stress)
            results: Dict[str, Any] = {"energy": energy, "forces": forces}
            if want_stress:
                if virial is None:
                    results["stress"] = np.zeros(6)
                else:
                    # ASE convention: stress = (1/V) dE/d(strain), Voigt order xx yy zz yz xz xy
                    sig = 0.5 * (virial + virial.T) / volume
                    results["stress"] = np.array([sig[0, 0], sig[1, 1], sig[2, 2], sig[1, 2], sig[0, 2], sig[0, 1]])
            self.results = results
            self._n_calls += 1
            if self.enable_timing:
                elapsed = time.perf_counter() - start
                self._total_time += elapsed
                self._call_times.append(elapsed)
        except ValueError:
            raise
        except Exception as e:
            logger.error("Calculation failed: %s", e, exc_info=True)
            raise RuntimeError(f"Failed to calculate properties for {len(atoms)} atoms: {e}") from e

    def _validate_inputs(self, positions, numbers, cell, pbc):
        """Same checks and messages as ase_calculator.py:434-459; Z is validated against the
        embedding size (the reference checks 1-118 although trained tables stop at max_z)."""
        if len(positions) == 0:
            raise ValueError("Cannot calculate properties for empty structure")
        z_min, z_max = int(numbers.min()), int(numbers.max())   # two reductions instead of three masks: this runs every MD step
        if z_min < 1 or z_max > 118:
            raise ValueError(f"Invalid atomic numbers: must be 1-118, got {numbers}")
        if z_max > self.model.max_z:
            raise ValueError(f"Invalid atomic numbers: model supports Z <= {self.model.max_z}, got {numbers}")
        if not np.isfinite(positions).all():
            raise ValueError("Positions contain NaN or Inf values")
        if pbc.any():
            if not np.isfinite(cell).all():
                raise ValueError("Cell contains NaN or Inf values")
            volume = abs(np.linalg.det(cell))
            if volume < 1e-6:
                warnings.warn(f"Cell volume very small ({volume:.2e} Å³), may indicate degenerate cell",
                              stacklevel=3)
            if self.pbc_mode == "minimum_image":
                self._check_minimum_image(cell, pbc)

    def _check_minimum_image(self, cell, pbc):
        """Minimum image is unique only if every periodic cell height is >= 2 (r_c + skin)."""
        check_minimum_image(cell, pbc, float(self.model.cutoff), getattr(self, "skin", 0.0))

    def _evaluate_single(self, positions, numbers, cell, pbc, want_virial: bool = False):
        """One structure.  The first calls for a system run eagerly (size the workspace, grow it
        on overflow); from the fourth call on the whole step -- pinned H2D of the positions, the
        ~30 kernels, status words and D2H of the results -- is ONE CUDA-graph replay: launching
        the kernels one by one from Python costs more host time (~110 us) than a small system
        needs on the device (92 us for H2O).  ``use_graph=False`` keeps every call eager."""
        dev = self.device
        n = len(numbers)
        periodic = self.pbc_mode == "minimum_image" and bool(pbc.any())
        cell_key = np.asarray(cell, dtype=np.float64).tobytes() + np.asarray(pbc, dtype=bool).tobytes() if periodic else b""
        c = self._numbers_cache
        if (c is None or len(c["numbers"]) != n or not np.array_equal(c["numbers"], numbers)
                or c["cell_key"] != cell_key):
            cells_d = pbc_d = None
            if periodic:
                cells_d, pbc_d = StudentForceField.pack_cells(torch.from_numpy(np.asarray(cell)), torch.from_numpy(np.asarray(pbc)), 1, dev)
            c = self._numbers_cache = {
                "numbers": np.array(numbers), "cell_key": cell_key,
                "z_d": torch.from_numpy(np.ascontiguousarray(numbers, dtype=np.int32)).to(dev),
                "off_d": torch.tensor([0, n], dtype=torch.int32, device=dev),
                "pos_d": torch.empty((n, 3), dtype=torch.float32, device=dev),
                "e_d": torch.empty(1, dtype=torch.float32, device=dev),
                "f_d": torch.empty((n, 3), dtype=torch.float32, device=dev),
                "w_d": torch.empty((1, 3, 3), dtype=torch.float32, device=dev),
                "cells_d": cells_d, "pbc_d": pbc_d,
                "pin_pos": torch.empty((n, 3), dtype=torch.float32).pin_memory(),
                "pin_out": torch.empty(3 * n + 10, dtype=torch.float32).pin_memory(),
                "pin_status": torch.zeros(6, dtype=torch.int32).pin_memory(),
                "graphs": {}, "warm": False, "graph_key": None, "calls": 0,
            }
            c["status_np"], c["out_np"], c["pin_pos_np"] = c["pin_status"].numpy(), c["pin_out"].numpy(), c["pin_pos"].numpy()
        c["pin_pos_np"][...] = positions   # FP64 -> FP32 conversion straight into the pinned buffer
        eng = self.model.engine()
        stream = torch.cuda.current_stream(dev)

        def enqueue_step():
            c["pos_d"].copy_(c["pin_pos"], non_blocking=True)
            eng.energy_forces_async(c["z_d"], c["pos_d"], c["off_d"], 1, c["e_d"], c["f_d"], c["cells_d"], c["pbc_d"])
            if want_virial:
                eng.virial_async(c["off_d"], 1, c["w_d"])
            eng.status_async(c["pin_status"])
            c["pin_out"][0:1].copy_(c["e_d"], non_blocking=True)
            c["pin_out"][1:3 * n + 1].copy_(c["f_d"].view(-1), non_blocking=True)
            if want_virial:
                c["pin_out"][3 * n + 1:3 * n + 10].copy_(c["w_d"].view(-1), non_blocking=True)

        graph_key = (id(eng), eng.cap_atoms, eng.cap_edges, eng.cap_structs)
        if c["graph_key"] != graph_key:       # workspace (re)allocated: captured pointers are stale
            c["graphs"], c["graph_key"] = {}, graph_key
        done = False
        c["calls"] += 1
        # capturing costs tens of milliseconds: only for a system that keeps coming back (MD, relaxation),
        # not for a screening loop that sees every structure once or twice
        if (self.use_graph and c["warm"] and c["calls"] > self.graph_after_calls and c["graph_key"] == graph_key
                and not eng.profiling):
            g = c["graphs"].get(want_virial)
            if g is None:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(stream)
                with torch.cuda.stream(side):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        enqueue_step()
                stream.wait_stream(side)
                c["graphs"][want_virial] = g
            g.replay()
            stream.synchronize()
            done = not (c["status_np"][2] or c["status_np"][5])   # overflow / FP16 saturation: eager path
        if not done:
            eng.ensure(n, 1, self.model._edges_per_atom)
            for _ in range(3):
                enqueue_step()
                stream.synchronize()
                if c["status_np"][5] and not c["status_np"][2]:
                    # a tensor-core operand left the FP16 range: this call again on the FP32 kernels
                    eng.set_dense_fallback(True)
                    try:
                        enqueue_step()
                        stream.synchronize()
                    finally:
                        eng.set_dense_fallback(False)
                    eng.saturation_reruns += 1
                if not c["status_np"][2]:
                    break
                num_edges = int(c["status_np"][0])
                eng.reserve(n, int(num_edges * 1.25) + 64, 1)
                self.model._edges_per_atom = max(self.model._edges_per_atom, int(num_edges * 1.25 / max(n, 1)) + 1)
            else:
                raise RuntimeError("edge workspace overflow persisted after growing")
            c["warm"] = True
        out = c["out_np"]
        virial = out[3 * n + 1:3 * n + 10].reshape(3, 3).astype(np.float64) if want_virial else None
        return float(out[0]), out[1:3 * n + 1].reshape(n, 3).copy(), virial

    # ---- batches --------------------------------------------------------------------------
    def calculate_batch(self, atoms_list, properties: Sequence[str] = ("energy", "forces")
                        ) -> List[Dict[str, Any]]:
        """One fused evaluation for many structures (ase_calculator.py:590-645).  ``[]`` -> ``[]``;
        a single structure takes the single path; PBC is ignored in batch mode like the
        reference (:735-736) unless ``pbc_mode='minimum_image'``."""
        if not atoms_list:
            return []
        if len(atoms_list) == 1:
            atoms = atoms_list[0]
            atoms.calc = self
            return [{   # same keys and stress handling as the reference's single-structure branch (:626-633)
                "energy": self.get_potential_energy(atoms) if "energy" in properties else None,
                "forces": self.get_forces(atoms) if "forces" in properties else None,
                "stress": self.get_stress(atoms) if "stress" in properties and self.enable_stress else None,
            }]
        counts = np.fromiter(map(len, atoms_list), dtype=np.int64, count=len(atoms_list))
        if np.any(counts == 0):
            raise ValueError("Cannot calculate properties for empty structure")
        # marshalling (ase_calculator.py:647-706 stacks the same two arrays): ASE's `numbers` / `positions`
        # attributes are the stored arrays, the getters copy each of them first -- 1024 small copies per call
        try:
            numbers = np.concatenate([a.numbers for a in atoms_list])
            positions = np.concatenate([a.positions for a in atoms_list])
        except AttributeError:   # Atoms-like objects that only offer the getters
            numbers = np.concatenate([a.get_atomic_numbers() for a in atoms_list])
            positions = np.concatenate([a.get_positions() for a in atoms_list])
        cells = pbcs = None
        if self.pbc_mode == "minimum_image" and any(np.any(a.get_pbc()) for a in atoms_list):
            cells = np.stack([np.asarray(a.get_cell(), dtype=np.float64) for a in atoms_list])
            pbcs = np.stack([np.asarray(a.get_pbc(), dtype=bool) for a in atoms_list])
        energies, forces = self.evaluate_arrays(numbers, positions, counts, cells, pbcs)
        # per-structure result dicts (ase_calculator.py:765-817): same keys, forces are views of one array
        want_e, want_f = "energy" in properties, "forces" in properties
        want_s = "stress" in properties and self.enable_stress   # "not supported in batch mode yet" (:810-812)
        e_list = energies.tolist()
        offs = np.concatenate([[0], np.cumsum(counts)]).tolist()
        results = []
        for i in range(len(counts)):
            r: Dict[str, Any] = {}
            if want_e:
                r["energy"] = e_list[i]
            if want_f:
                r["forces"] = forces[offs[i]:offs[i + 1]]
            if want_s:
                r["stress"] = None
            results.append(r)
        return results

    def evaluate_arrays(self, numbers: np.ndarray, positions: np.ndarray, counts: np.ndarray,
                        cells: Optional[np.ndarray] = None, pbcs: Optional[np.ndarray] = None,
                        out: Optional[Tuple[np.ndarray, np.ndarray]] = None):
        """Host arrays in, host arrays out: (energies [B] float32, forces [N,3] float32).  The
        batched-structure interface underneath ``calculate_batch``; validation included.
        ``out = (energies, forces)``: caller-provided float32 arrays of those shapes (e.g. a rank's slice
        of a shared-memory result, ``sharding.SharedResults``) that receive the results of every
        micro-batch as it completes; they are also what is returned."""
        if out is not None:
            e_out, f_out = out
            if e_out.shape != (len(counts),) or f_out.shape != (len(numbers), 3) \
                    or e_out.dtype != np.float32 or f_out.dtype != np.float32:
                raise ValueError("out must be (float32 [structures], float32 [atoms, 3]) arrays")
            e, f = self._evaluate_arrays(numbers, positions, counts, cells, pbcs, out)
            if e is not e_out:   # paths that produce their own arrays
                e_out[...] = e
                f_out[...] = f
            return e_out, f_out
        return self._evaluate_arrays(numbers, positions, counts, cells, pbcs, None)

    def _evaluate_arrays(self, numbers, positions, counts, cells, pbcs, out):
        numbers = np.asarray(numbers)
        positions = np.asarray(positions)
        counts = np.asarray(counts, dtype=np.int64)
        self._validate_arrays(numbers, positions, counts)
        periodic = cells is not None and pbcs is not None and bool(np.any(pbcs))
        if periodic:
            if not np.isfinite(np.asarray(cells)).all():
                raise ValueError("Cell contains NaN or Inf values")
            # a cell too small for the minimum image would silently drop periodic images
            self._check_minimum_image(np.asarray(cells, dtype=np.float64), np.asarray(pbcs, dtype=bool))
        if self._replicas is not None and not periodic and len(counts) >= 2 * len(self._replicas):
            return self._evaluate_on_all_devices(numbers, positions, counts)
        if len(numbers) > self.max_atoms_per_call and len(counts) > 1:
            # micro-batches that fit the workspace (a 100 k-structure sweep does not fit in one call)
            from .sharding import chunk_by_budget
            offs = np.concatenate([[0], np.cumsum(counts)])
            chunks = chunk_by_budget(counts, self.max_atoms_per_call, 1 << 20)
            e_parts, f_parts = [], []
            if cells is None or pbcs is None or not np.any(pbcs):
                # open boundaries: keep the copies of chunk k+1 / k-1 under the kernels of chunk k
                gen = ((numbers[int(offs[a]):int(offs[b])], positions[int(offs[a]):int(offs[b])], counts[a:b])
                       for a, b in chunks)
                if out is not None:   # results of chunk k land in the caller's arrays while chunk k+1 runs
                    for (a, b), (e, f) in zip(chunks, self.evaluate_stream(gen)):
                        out[0][a:b] = e
                        out[1][int(offs[a]):int(offs[b])] = f
                    return out
                for e, f in self.evaluate_stream(gen):
                    e_parts.append(e)
                    f_parts.append(f)
            else:
                for a, b in chunks:
                    sl = slice(int(offs[a]), int(offs[b]))
                    e, f = self.evaluate_arrays(numbers[sl], positions[sl], counts[a:b], cells[a:b], pbcs[a:b])
                    e_parts.append(e)
                    f_parts.append(f)
            return np.concatenate(e_parts), np.concatenate(f_parts)
        if cells is None or pbcs is None or not np.any(pbcs):
            # open boundaries: one batch through the pipelined interface's slot machinery -- copies on
            # the copy streams, status words behind the step, ONE host synchronisation
            try:
                return next(iter(self.evaluate_stream([(numbers, positions, counts)])))
            except ValueError:
                raise
            except Exception as e:
                raise RuntimeError(f"Failed to calculate properties for {len(numbers)} atoms: {e}") from e
        dev = self.device
        nb, n = len(counts), len(numbers)
        st = self._batch_staging(n, nb)
        # host -> pinned staging (dtype conversion happens in this copy) -> device, async
        np.copyto(st["z_h"].numpy()[:n], numbers, casting="unsafe")
        st["pos_h"][:n].copy_(torch.from_numpy(np.ascontiguousarray(positions)))
        off_np = st["off_h"].numpy()
        off_np[0] = 0
        np.cumsum(counts, out=off_np[1:nb + 1])
        z_d, pos_d, off_d = st["z_d"][:n], st["pos_d"][:n], st["off_d"][:nb + 1]
        z_d.copy_(st["z_h"][:n], non_blocking=True)
        pos_d.copy_(st["pos_h"][:n], non_blocking=True)
        off_d.copy_(st["off_h"][:nb + 1], non_blocking=True)
        cells_d = pbc_d = None
        if cells is not None and pbcs is not None and np.any(pbcs):
            cells_d, pbc_d = StudentForceField.pack_cells(torch.from_numpy(np.asarray(cells)),
                                                          torch.from_numpy(np.asarray(pbcs)), nb, dev)
        try:
            e_d, f_d = self.model.energy_and_forces_packed(z_d, pos_d, off_d, nb, cells_d, pbc_d,
                                                           max_atoms=int(counts.max()))
            st["e_h"][:nb].copy_(e_d, non_blocking=True)
            st["f_h"][:n].copy_(f_d, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            energies = st["e_h"][:nb].numpy().copy()
            forces = st["f_h"][:n].numpy().copy()
        except Exception as e:
            raise RuntimeError(f"Failed to calculate properties for {len(numbers)} atoms: {e}") from e
        self._n_calls += 1
        return energies, forces

    def _evaluate_on_all_devices(self, numbers, positions, counts):
        """One structure list over ``device_ids``: contiguous shards balanced by atoms
        (sharding.partition_by_atoms), one host thread per device driving that device's own pipeline
        (pinned staging, copy streams, micro-batches), results concatenated in input order."""
        from .sharding import partition_by_atoms
        offs = np.concatenate([[0], np.cumsum(counts)])
        shards = partition_by_atoms(counts, len(self._replicas))

        def work(replica, a, b):
            if b <= a:
                return np.zeros(0, dtype=np.float32), np.zeros((0, 3), dtype=np.float32)
            torch.cuda.set_device(replica.device)
            sl = slice(int(offs[a]), int(offs[b]))
            saved, replica._replicas = replica._replicas, None      # the replica evaluates its shard locally
            try:
                return replica.evaluate_arrays(numbers[sl], positions[sl], counts[a:b])
            finally:
                replica._replicas = saved

        futures = [self._pool.submit(work, r, a, b) for r, (a, b) in zip(self._replicas, shards)]
        parts = [f.result() for f in futures]
        self._n_calls += 1
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    # ---- pipelined batches -----------------------------------------------------------------
    def evaluate_stream(self, batches: Iterable[Tuple[np.ndarray, np.ndarray, np.ndarray]]
                        ) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
        """Screening-sweep form of :meth:`evaluate_arrays`: takes an iterable of
        ``(numbers, positions, counts)`` host batches (open boundaries) and yields
        ``(energies [B] float32, forces [N,3] float32)`` per batch, in order.

        Same work per batch as ``evaluate_arrays`` -- validation, host -> pinned -> device copy of
        the inputs, one fused energy+force step, device -> pinned -> host copy of the results --
        but two batches are in flight: while the kernels of batch k run on the compute stream, the
        host converts and uploads batch k+1 on a copy stream and the results of batch k-1 come
        back on another, so the GPU never waits for the host.  The step's status words travel
        with the results (``mlffd_status_async``): an edge-workspace overflow is detected when the
        batch is collected and that batch is re-run through the blocking path."""
        dev = self.device
        eng = self.model.engine()
        compute = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_streams", None) is None:
            self._copy_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            self._slots = [{"cap_n": 0, "cap_b": 0} for _ in range(2)]
        s_in, s_out = self._copy_streams

        def prepare(batch, slot):
            numbers, positions, counts = (np.asarray(batch[0]), np.asarray(batch[1]),
                                          np.asarray(batch[2], dtype=np.int64))
            self._validate_arrays(numbers, positions, counts)
            nb, n = len(counts), len(numbers)
            if n > self.max_atoms_per_call and nb > 1:
                raise ValueError(f"evaluate_stream: a batch holds {n} atoms; keep batches under "
                                 f"max_atoms_per_call = {self.max_atoms_per_call}")
            self._grow_slot(slot, n, nb)
            slot["n"], slot["nb"], slot["max_count"] = n, nb, int(counts.max())
            # int64 -> int32 with numpy (13 us for 51 200 atoms): torch's converting copy_ takes its thread pool for
            # this, which costs milliseconds on a host whose cores are busy with other ranks
            np.copyto(slot["z_h"].numpy()[:n], numbers, casting="unsafe")
            slot["pos_h"][:n].copy_(torch.from_numpy(np.ascontiguousarray(positions)))
            off_np = slot["off_h"].numpy()
            off_np[0] = 0
            np.cumsum(counts, out=off_np[1:nb + 1])
            with torch.cuda.stream(s_in):
                if slot.get("compute_done") is not None:   # the slot's previous tenant has been read
                    s_in.wait_event(slot["compute_done"])
                slot["z_d"][:n].copy_(slot["z_h"][:n], non_blocking=True)
                slot["pos_d"][:n].copy_(slot["pos_h"][:n], non_blocking=True)
                slot["off_d"][:nb + 1].copy_(slot["off_h"][:nb + 1], non_blocking=True)
                slot["h2d_done"] = s_in.record_event()
            return slot

        def launch(slot):
            n, nb = slot["n"], slot["nb"]
            compute.wait_event(slot["h2d_done"])
            eng.ensure(n, nb, self.model._edges_per_atom)
            eng.energy_forces_async(slot["z_d"][:n], slot["pos_d"][:n], slot["off_d"][:nb + 1], nb,
                                    slot["e_d"][:nb], slot["f_d"][:n])
            eng.status_async(slot["status_h"])   # pinned host words, ordered behind the step
            slot["compute_done"] = compute.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(slot["compute_done"])
                slot["e_h"][:nb].copy_(slot["e_d"][:nb], non_blocking=True)
                slot["f_h"][:n].copy_(slot["f_d"][:n], non_blocking=True)
                slot["out_done"] = s_out.record_event()

        def collect(slot):
            n, nb = slot["n"], slot["nb"]
            slot["out_done"].synchronize()
            status = slot["status_h"].numpy()
            if status[2] or status[5]:   # overflow / FP16 saturation: the blocking path grows / falls back and re-runs
                torch.cuda.synchronize(dev)
                try:
                    e_d, f_d = self.model.energy_and_forces_packed(
                        slot["z_d"][:n], slot["pos_d"][:n], slot["off_d"][:nb + 1], nb,
                        max_atoms=slot["max_count"])
                    energies, forces = e_d.cpu().numpy(), f_d.cpu().numpy()
                except Exception as e:
                    raise RuntimeError(f"Failed to calculate properties for {n} atoms: {e}") from e
            else:
                energies = slot["e_h"][:nb].numpy().copy()
                forces = slot["f_h"][:n].numpy().copy()
            self._n_calls += 1
            return energies, forces

        it = iter(batches)
        first = next(it, None)
        if first is None:
            return
        k, prev = 0, None
        cur = prepare(first, self._slots[0])
        while cur is not None:
            launch(cur)
            if prev is not None:
                yield collect(prev)
            nxt_batch = next(it, None)
            k += 1
            nxt = prepare(nxt_batch, self._slots[k % 2]) if nxt_batch is not None else None
            prev, cur = cur, nxt
        yield collect(prev)

    def _validate_arrays(self, numbers, positions, counts):
        if len(numbers) == 0 or len(counts) == 0 or np.any(counts == 0):
            raise ValueError("Cannot calculate properties for empty structure")
        if int(counts.sum()) != len(numbers) or len(positions) != len(numbers):
            raise ValueError("counts / numbers / positions disagree on the number of atoms")
        if int(numbers.min()) < 1 or int(numbers.max()) > min(118, self.model.max_z):
            raise ValueError(f"Invalid atomic numbers: must be 1-{min(118, self.model.max_z)}")
        if not np.isfinite(positions).all():
            raise ValueError("Positions contain NaN or Inf values")

    def _grow_slot(self, slot, n_atoms: int, n_structs: int):
        """Grow-only pinned + device buffers of one pipeline slot."""
        if slot["cap_n"] >= n_atoms and slot["cap_b"] >= n_structs:
            return
        dev = self.device
        torch.cuda.synchronize(dev)   # nothing may still read the buffers being replaced
        cap_n = max(n_atoms, int(1.25 * slot["cap_n"]))
        cap_b = max(n_structs, int(1.25 * slot["cap_b"]))
        slot.update({
            "cap_n": cap_n, "cap_b": cap_b,
            "z_h": torch.empty(cap_n, dtype=torch.int32).pin_memory(),
            "pos_h": torch.empty((cap_n, 3), dtype=torch.float32).pin_memory(),
            "off_h": torch.empty(cap_b + 1, dtype=torch.int32).pin_memory(),
            "e_h": torch.empty(cap_b, dtype=torch.float32).pin_memory(),
            "f_h": torch.empty((cap_n, 3), dtype=torch.float32).pin_memory(),
            "status_h": torch.zeros(6, dtype=torch.int32).pin_memory(),
            "z_d": torch.empty(cap_n, dtype=torch.int32, device=dev),
            "pos_d": torch.empty((cap_n, 3), dtype=torch.float32, device=dev),
            "off_d": torch.empty(cap_b + 1, dtype=torch.int32, device=dev),
            "e_d": torch.empty(cap_b, dtype=torch.float32, device=dev),
            "f_d": torch.empty((cap_n, 3), dtype=torch.float32, device=dev),
            "compute_done": None,
        })

    def _batch_staging(self, n_atoms: int, n_structs: int):
        """Grow-only pinned host + device staging buffers for the batched interface."""
        st = getattr(self, "_staging", None)
        if st is None or st["cap_n"] < n_atoms or st["cap_b"] < n_structs:
            cap_n = max(n_atoms, int(1.25 * (st["cap_n"] if st else 0)))
            cap_b = max(n_structs, int(1.25 * (st["cap_b"] if st else 0)))
            dev = self.device
            st = {
                "cap_n": cap_n, "cap_b": cap_b,
                "z_h": torch.empty(cap_n, dtype=torch.int32).pin_memory(),
                "pos_h": torch.empty((cap_n, 3), dtype=torch.float32).pin_memory(),
                "off_h": torch.empty(cap_b + 1, dtype=torch.int32).pin_memory(),
                "e_h": torch.empty(cap_b, dtype=torch.float32).pin_memory(),
                "f_h": torch.empty((cap_n, 3), dtype=torch.float32).pin_memory(),
                "z_d": torch.empty(cap_n, dtype=torch.int32, device=dev),
                "pos_d": torch.empty((cap_n, 3), dtype=torch.float32, device=dev),
                "off_d": torch.empty(cap_b + 1, dtype=torch.int32, device=dev),
            }
            self._staging = st
        return st

    # ---- bookkeeping ----------------------------------------------------------------------
    def reset(self):
        _AseCalculator.reset(self)
        self._numbers_cache = None

    @property
    def n_calls(self) -> int:
        return self._n_calls

    @property
    def avg_time(self) -> float:
        return 0.0 if self._n_calls == 0 else self._total_time / self._n_calls

    def get_timing_stats(self) -> Dict[str, float]:
        if not self._call_times:
            return {"n_calls": 0, "total_time": 0.0, "avg_time": 0.0, "min_time": 0.0,
                    "max_time": 0.0, "median_time": 0.0}
        t = np.array(self._call_times)
        return {"n_calls": self._n_calls, "total_time": self._total_time,
                "avg_time": float(np.mean(t)), "min_time": float(np.min(t)),
                "max_time": float(np.max(t)), "median_time": float(np.median(t)),
                "std_time": float(np.std(t))}

    def __repr__(self) -> str:
        return (f"StudentForceFieldCalculator(checkpoint={self.checkpoint_path.name}, "
                f"device={self.device}, calls={self._n_calls})")


__all__ = ["StudentForceFieldCalculator"]
