"""Host-side NVE driver used by configs C1/C3/C4 when ASE is not installed.

Restates what the reference's MD harness does around the calculator
(src/mlff_distiller/testing/nve_harness.py:141-171 initial velocities, :214-235 VelocityVerlet,
:329-331 drift metric; energy_metrics.py:74-80): Maxwell-Boltzmann velocities at T0 with the
centre-of-mass translation (and optionally rotation) removed, textbook velocity-Verlet
half-kick / drift / half-kick in ASE units (eV, Angstrom, amu; time unit
``fs = 1e-15 s * sqrt(e/amu) / Angstrom``), and ``drift % = 100 (E_tot[-1] - E_tot[0]) / |E_tot[0]|``
with E_tot = PE + KE sampled every step.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np

# CODATA-2014 constants as ASE uses them (ase/units.py): eV, Angstrom, amu base units.
_E = 1.6021766208e-19
_AMU = 1.660539040e-27
KB = 8.6173303e-5  # eV / K
FS = 1e-15 * np.sqrt(_E / _AMU) * 1e10  # = 0.09822694788...

# Standard atomic weights (IUPAC abridged), index = atomic number.
ATOMIC_MASSES = np.array([
    1.0, 1.008, 4.002602, 6.94, 9.0121831, 10.81, 12.011, 14.007, 15.999, 18.998403163, 20.1797,
    22.98976928, 24.305, 26.9815385, 28.085, 30.973761998, 32.06, 35.45, 39.948, 39.0983, 40.078,
    44.955908, 47.867, 50.9415, 51.9961, 54.938044, 55.845, 58.933194, 58.6934, 63.546, 65.38,
    69.723, 72.630, 74.921595, 78.971, 79.904, 83.798, 85.4678, 87.62, 88.90584, 91.224,
    92.90637, 95.95, 97.90721, 101.07, 102.90550, 106.42, 107.8682, 112.414, 114.818, 118.710,
    121.760, 127.60, 126.90447, 131.293, 132.90545196, 137.327, 138.90547, 140.116, 140.90766,
    144.242, 144.91276, 150.36, 151.964, 157.25, 158.92535, 162.500, 164.93033, 167.259,
    168.93422, 173.054, 174.9668, 178.49, 180.94788, 183.84, 186.207, 190.23, 192.217, 195.084,
    196.966569, 200.592, 204.38, 207.2, 208.98040, 208.98243, 209.98715, 222.01758, 223.01974,
    226.02541, 227.02775, 232.0377, 231.03588, 238.02891, 237.04817, 244.06421, 243.06138,
    247.07035, 247.07031, 251.07959, 252.0830, 257.09511, 258.09843, 259.1010, 262.110, 267.122,
    268.126, 271.134, 270.133, 269.1338, 278.156, 281.165, 281.166, 285.177, 286.182, 289.190,
    289.194, 293.204, 293.208, 294.214])


def maxwell_boltzmann(masses: np.ndarray, temperature_K: float, rng: np.random.Generator,
                      positions: Optional[np.ndarray] = None, zero_rotation: bool = False
                      ) -> np.ndarray:
    """Velocities [N,3] (ASE units) drawn at T, COM momentum removed (nve_harness.py:158-165)."""
    m = np.asarray(masses, dtype=np.float64)[:, None]
    v = rng.normal(size=(len(m), 3)) * np.sqrt(KB * temperature_K / m)
    v -= (m * v).sum(0) / m.sum()
    if zero_rotation and positions is not None and len(m) > 2:
        x = positions - (m * positions).sum(0) / m.sum()
        L = (m * np.cross(x, v)).sum(0)
        I = np.zeros((3, 3))
        for a in range(len(m)):
            r = x[a]
            I += m[a, 0] * ((r @ r) * np.eye(3) - np.outer(r, r))
        try:
            omega = np.linalg.solve(I, L)
            v -= np.cross(omega, x)
        except np.linalg.LinAlgError:
            pass
    return v


def kinetic_energy(masses: np.ndarray, velocities: np.ndarray) -> float:
    return float(0.5 * np.sum(np.asarray(masses)[:, None] * velocities ** 2))


def energy_drift_percent(e_total: np.ndarray) -> float:
    """nve_harness.py:329-331."""
    return float(100.0 * (e_total[-1] - e_total[0]) / abs(e_total[0]))


def velocity_verlet(force_fn: Callable[[np.ndarray], Tuple[float, np.ndarray]],
                    positions: np.ndarray, velocities: np.ndarray, masses: np.ndarray,
                    steps: int, dt_fs: float = 0.5, record_every: int = 1) -> Dict[str, np.ndarray]:
    """NVE trajectory; ``force_fn(positions) -> (potential energy, forces[N,3])``.

    One force evaluation per step, float64 integrator on the host, like ASE's VelocityVerlet.
    Returns positions/velocities at the end and the PE/KE/E_tot series (step 0 included).
    """
    dt = dt_fs * FS
    x = np.array(positions, dtype=np.float64)
    v = np.array(velocities, dtype=np.float64)
    m = np.asarray(masses, dtype=np.float64)[:, None]
    pe, f = force_fn(x)
    f = np.asarray(f, dtype=np.float64)
    pes, kes = [float(pe)], [kinetic_energy(m[:, 0], v)]
    for step in range(1, steps + 1):
        v += 0.5 * dt * f / m
        x += dt * v
        pe, f = force_fn(x)
        f = np.asarray(f, dtype=np.float64)
        v += 0.5 * dt * f / m
        if step % record_every == 0 or step == steps:
            pes.append(float(pe))
            kes.append(kinetic_energy(m[:, 0], v))
    pes_a, kes_a = np.asarray(pes), np.asarray(kes)
    return {"positions": x, "velocities": v, "potential": pes_a, "kinetic": kes_a,
            "total": pes_a + kes_a, "drift_percent": energy_drift_percent(pes_a + kes_a)}


def ns_per_day(steps_per_second: float, dt_fs: float = 0.5) -> float:
    """ns/day = steps/s * dt[fs] * 86400 * 1e-6."""
    return steps_per_second * dt_fs * 86400.0 * 1e-6


class DeviceMD:
    """NVE trajectory with positions, velocities and forces resident on the GPU.

    One step = ``mlffd_md_kick_drift`` -> ``mlffd_energy_forces`` -> ``mlffd_md_kick_energy``
    (include/mlffd.h), optionally captured once into a CUDA graph and replayed, so a step costs one
    graph launch and no host<->device copy; the (PE, KE) series is read back after ``run``.
    Integrator state is FP64 on the device; the model sees FP32 positions, like the reference's
    calculator does every step.  ``structures`` may hold several independent systems (offsets).
    """

    def __init__(self, model, numbers, positions, velocities, masses, dt_fs: float = 0.5,
                 offsets=None, cell=None, pbc=None, use_graph: bool = True, capacity: int = 1 << 20,
                 edge_reserve: Optional[int] = None):
        import ctypes
        import torch
        self.torch, self.ctypes = torch, ctypes
        self.model, self.eng = model, model.engine()
        self.lib = self.eng.lib
        dev = self.eng.device
        n = len(numbers)
        self.n, self.dt = n, float(dt_fs) * FS
        offsets = np.array([0, n], dtype=np.int32) if offsets is None else np.asarray(offsets, dtype=np.int32)
        self.nb = len(offsets) - 1
        f64 = dict(dtype=torch.float64, device=dev)
        self.z = torch.from_numpy(np.ascontiguousarray(numbers, dtype=np.int32)).to(dev)
        self.off = torch.from_numpy(offsets).to(dev)
        self.pos = torch.tensor(np.asarray(positions, dtype=np.float64).reshape(n, 3), **f64)
        self.vel = torch.tensor(np.asarray(velocities, dtype=np.float64).reshape(n, 3), **f64)
        self.inv_mass = torch.tensor(1.0 / np.asarray(masses, dtype=np.float64), **f64)
        self.pos32 = self.pos.to(torch.float32)
        self.energy = torch.zeros(self.nb, dtype=torch.float32, device=dev)
        self.forces = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        self.capacity = int(capacity)
        self.series = torch.zeros((self.capacity, 2), **f64)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cells = self.pbc = None
        if cell is not None and pbc is not None and bool(np.any(pbc)):
            from .student_model import check_minimum_image
            check_minimum_image(np.asarray(cell), np.asarray(pbc), float(model.cutoff), float(getattr(model, "skin", 0.0)))
            self.cells, self.pbc = model.pack_cells(torch.as_tensor(np.asarray(cell)), torch.as_tensor(np.asarray(pbc)),
                                                    self.nb, dev)
        self.use_graph, self.graph, self._graph_caps = use_graph, None, None
        # forces at t = 0 (sizes the workspace; 1.5x edge head-room for density fluctuations)
        for _ in range(3):
            self.eng.ensure(n, self.nb)
            self._forces()
            st = self.eng.status()
            if not st.overflow:
                # edge_reserve: explicit capacity (tests of the freeze / resume path); default 1.5x head-room
                self.eng.reserve(n, int(st.num_edges * 1.5) + 64 if edge_reserve is None else int(edge_reserve), self.nb)
                break
            self.eng.reserve(n, int(st.num_edges * 1.5) + 64, self.nb)
        self._forces()
        self.interruptions = 0   # steps that had to be repaired (workspace growth / FP32 fallback) and resumed
        self._record(0.0)  # series[0] = (PE, KE) at t = 0

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def _forces(self):
        self.eng.energy_forces_async(self.z, self.pos32, self.off, self.nb, self.energy, self.forces,
                                     self.cells, self.pbc)

    def _record(self, dt):
        rc = self.lib.mlffd_md_kick_energy(self.eng._ctx, self.n, self.vel.data_ptr(), self.forces.data_ptr(),
                                           self.inv_mass.data_ptr(), dt, self.energy.data_ptr(), self.nb,
                                           self.series.data_ptr(), self.counter.data_ptr(), self.capacity,
                                           self.eng._stream())
        if rc:
            raise RuntimeError(f"mlffd_md_kick_energy failed ({rc})")

    def _step(self):
        rc = self.lib.mlffd_md_kick_drift(self.eng._ctx, self.n, self.pos.data_ptr(), self.vel.data_ptr(), self.forces.data_ptr(),
                                          self.inv_mass.data_ptr(), self.dt, self.pos32.data_ptr(), self.eng._stream())
        if rc:
            raise RuntimeError(f"mlffd_md_kick_drift failed ({rc})")
        self._forces()
        self._record(self.dt)

    def _capture(self):
        """Capture one step as a CUDA graph (warm-up outside the capture; the trajectory does not advance)."""
        torch = self.torch
        side = torch.cuda.Stream(device=self.eng.device)
        side.wait_stream(torch.cuda.current_stream(self.eng.device))
        with torch.cuda.stream(side):
            state = (self.pos.clone(), self.vel.clone(), self.forces.clone(), self.energy.clone(),
                     self.counter.clone())

            def restore():
                for dst, src in zip((self.pos, self.vel, self.forces, self.energy, self.counter), state):
                    dst.copy_(src)
                self.pos32.copy_(self.pos.to(torch.float32))

            self._step()
            side.synchronize()
            restore()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._step()
            restore()   # the capture run itself must not advance the trajectory
            self._forces()   # device status words of the restored state (the kicks are guarded by them)
        torch.cuda.current_stream(self.eng.device).wait_stream(side)
        self._graph_caps = (self.eng.cap_atoms, self.eng.cap_edges, self.eng.cap_structs, self.eng.dense_fallback)

    def run(self, steps: int) -> Dict[str, np.ndarray]:
        """Advance ``steps`` steps; returns the energy series so far (step 0 included).

        Steps are enqueued without host synchronisation.  If a force evaluation fails on the device
        (edge-workspace overflow, FP16 saturation of a tensor-core operand) the guarded kick / drift
        kernels freeze the trajectory at that step's valid mid-step state; this method then grows the
        workspace (or moves the dense layers to the FP32 kernels), completes the interrupted step and
        carries on, so the returned series is the one an uninterrupted run would have produced."""
        torch = self.torch
        target = int(self.counter.item()) + int(steps)
        for _attempt in range(8):
            todo = target - int(self.counter.item())
            if todo <= 0:
                break
            caps = (self.eng.cap_atoms, self.eng.cap_edges, self.eng.cap_structs, self.eng.dense_fallback)
            if self.use_graph and (self.graph is None or self._graph_caps != caps):
                self._capture()
            for _ in range(todo):
                if self.use_graph:
                    self.graph.replay()
                else:
                    self._step()
            torch.cuda.synchronize(self.eng.device)
            if int(self.counter.item()) >= target:
                break
            # frozen at (x_k, v_{k-1/2}) of the failing step k: repair, finish step k, go round again
            st = self.eng.status()
            self.interruptions += 1
            if st.overflow:
                self.eng.reserve(self.n, int(st.num_edges * 1.5) + 64, self.nb)
            elif st.tc_saturated:
                self.eng.set_dense_fallback(True)
            self._forces()
            st = self.eng.status()
            if st.overflow or st.tc_saturated:
                continue
            self._record(self.dt)
        else:
            raise RuntimeError("on-device trajectory could not be resumed after repeated device-side failures")
        torch.cuda.synchronize(self.eng.device)
        out = self.energies()
        if not np.isfinite(out["total"]).all():
            raise RuntimeError("non-finite energies in the on-device trajectory")
        return out

    def energies(self) -> Dict[str, np.ndarray]:
        k = min(int(self.counter.item()), self.capacity)
        ser = self.series[:k].cpu().numpy()
        total = ser[:, 0] + ser[:, 1]
        return {"potential": ser[:, 0], "kinetic": ser[:, 1], "total": total,
                "drift_percent": energy_drift_percent(total) if k > 1 else 0.0}

    def state(self):
        return self.pos.cpu().numpy(), self.vel.cpu().numpy()
