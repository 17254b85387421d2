#!/usr/bin/env python
"""Generate the 10 ps NVE reference fixture (CPU, minutes): the CPU restatement of the reference path
(oracle/painn_oracle.py, FP32, the same ATen ops as src/mlff_distiller/models/student_model.py and
bit-identical to it on the golden cases) drives the float64 velocity-Verlet integrator for 20 000
steps of 0.5 fs from Maxwell-Boltzmann velocities at 300 K with the centre-of-mass translation and
rotation removed (reference src/mlff_distiller/testing/nve_harness.py:158-165, 214-235, 329-331).

    python tests/golden/make_nve_reference.py      # writes tests/golden/nve_reference.npz

Per trajectory the file holds the initial state (x0, v0), the total-energy series sub-sampled every
10 steps, the first 2 001 steps in full, and the end-point drift.  H2O: seeds 42..73 (BASELINE config
C1 uses seed 42); benzene: seed 42.  Six worker processes, ~12 minutes on 8 cores.  The GPU tests start on-device trajectories from exactly these
states and compare end-point drift and the early part of the series (tests/test_gpu_calculator.py).
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from mlff_distiller_b200 import md, synthetic  # noqa: E402
from oracle import painn_oracle as po  # noqa: E402

STEPS = 20000
H2O_SEEDS = range(42, 74)      # 32 trajectories: the end-point drift of H2O is a broad distribution


def run_job(job):
    name, seed = job
    torch.set_num_threads(1)   # 3 - 12 atoms: threading only adds overhead
    with np.load(ROOT / "tests" / "golden" / "weights_original.npz") as z:
        state = {k: z[k] for k in z.files if not k.startswith("__")}
        cutoff = json.loads(str(z["__config__"]))["cutoff"]
    w = po.to_torch_weights(state)
    atoms = synthetic.water() if name == "h2o" else synthetic.benzene()
    zt = torch.from_numpy(atoms.numbers)

    def ef(pos):
        e, f = po.energy_and_forces(w, zt, torch.from_numpy(pos.astype(np.float32)), cutoff, None)
        return float(e), f.numpy().astype(np.float64)

    m = atoms.get_masses()
    x0 = atoms.get_positions()
    v0 = md.maxwell_boltzmann(m, 300.0, np.random.default_rng(seed), x0, zero_rotation=True)
    t0 = time.time()
    traj = md.velocity_verlet(ef, x0, v0, m, steps=STEPS, dt_fs=0.5)
    key = f"{name}_seed{seed}"
    tot = traj["total"]
    out = {key + "_x0": x0, key + "_v0": v0, key + "_masses": m,
           key + "_total_every10": tot[::10], key + "_total_first2001": tot[:2001],
           key + "_drift_percent": traj["drift_percent"],
           key + "_band_percent": 100.0 * float(np.max(np.abs(tot - tot[0]))) / abs(tot[0])}
    print(f"{key}: drift {traj['drift_percent']:+.4f} %  band {out[key + '_band_percent']:.4f} %  "
          f"({time.time() - t0:.0f} s)", flush=True)
    return out


def main():
    import multiprocessing as mp
    jobs = [("benzene", 42)] + [("h2o", s) for s in H2O_SEEDS]
    out = {"steps": STEPS, "dt_fs": 0.5, "temperature_K": 300.0, "h2o_seeds": np.array(list(H2O_SEEDS))}
    with mp.get_context("spawn").Pool(max(1, min(6, (os.cpu_count() or 2) - 1))) as pool:
        for res in pool.imap_unordered(run_job, jobs):
            out.update(res)
    np.savez_compressed(ROOT / "tests" / "golden" / "nve_reference.npz", **out)


if __name__ == "__main__":
    main()
