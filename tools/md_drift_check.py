#!/usr/bin/env python
"""C3 drift diagnosis: same initial conditions, CUDA path vs the CPU oracle of the reference,
plus the dt-scaling of the drift (integrator error scales ~dt^2, force inconsistency does not)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mlff_distiller_b200 import md, synthetic  # noqa: E402
from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator  # noqa: E402
from oracle import painn_oracle as po  # noqa: E402

W = ROOT / "tests" / "golden"
with np.load(W / "weights_original.npz") as z:
    state = {k: z[k] for k in z.files if not k.startswith("__")}

chain = synthetic.alkane_chain(100)
chain.positions = chain.positions + np.random.default_rng(8).normal(0.0, 0.02, chain.positions.shape)
masses = chain.get_masses()
v0 = md.maxwell_boltzmann(masses, 300.0, np.random.default_rng(42), chain.positions, zero_rotation=True)
calc = StudentForceFieldCalculator(W / "weights_original.npz", device="cuda:0", precision=sys.argv[1] if len(sys.argv) > 1 else "tc")
work = chain.copy()


def gpu_force(x):
    work.set_positions(x)
    calc.calculate(work)
    return calc.results["energy"], calc.results["forces"]


def cpu_force(x):
    e, f = po.evaluate(state, 5.0, chain.numbers, x.astype(np.float32), [0, len(x)])
    return float(e[0]), f


out = {}
for dt in (0.5, 0.25, 0.125):
    steps = int(1000 * 0.5 / dt)
    r = md.velocity_verlet(gpu_force, chain.positions, v0, masses, steps, dt)
    out[f"gpu_dt{dt}"] = {"steps": steps, "drift_percent": r["drift_percent"], "dE": float(r["total"][-1] - r["total"][0]),
                          "KE_last": float(r["kinetic"][-1]), "PE_first": float(r["potential"][0]), "PE_last": float(r["potential"][-1]),
                          "series": [float(x) for x in r["total"][:: max(steps // 20, 1)]]}
    print(dt, json.dumps({k: v for k, v in out[f"gpu_dt{dt}"].items() if k != "series"}), flush=True)
torch.set_num_threads(16)
t0 = time.time()
r_cpu = md.velocity_verlet(cpu_force, chain.positions, v0, masses, 300, 0.5)
r_gpu = md.velocity_verlet(gpu_force, chain.positions, v0, masses, 300, 0.5)
out["cpu_300"] = {"drift_percent": r_cpu["drift_percent"], "dE": float(r_cpu["total"][-1] - r_cpu["total"][0]), "seconds": time.time() - t0}
out["gpu_300"] = {"drift_percent": r_gpu["drift_percent"], "dE": float(r_gpu["total"][-1] - r_gpu["total"][0])}
out["max_abs_dEtotal_gpu_vs_cpu_300"] = float(np.abs(r_cpu["total"] - r_gpu["total"]).max())
out["max_pos_diff_300"] = float(np.abs(r_cpu["positions"] - r_gpu["positions"]).max())
print(json.dumps({k: v for k, v in out.items() if k.startswith(("cpu", "gpu_300", "max"))}))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "md_drift_check.json").write_text(json.dumps(out, indent=1))
