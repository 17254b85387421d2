#!/usr/bin/env python
"""NVE energy conservation of the on-device Verlet for a small molecule: end-point drift (the
reference's metric, testing/energy_metrics.py:74-80), the amplitude of the total-energy oscillation
and the drift of a least-squares line through the series, for the current message-kernel settings.
Usage: [MLFFD_MSG_TEAM=0] python tools/md_drift_report.py [--systems h2o,benzene] [--steps 20000]"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from md_latency import system  # noqa: E402
from mlff_distiller_b200 import md  # noqa: E402
from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--systems", default="h2o,benzene")
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--seeds", default="42,43,44")
    ap.add_argument("--perturb", type=float, default=0.0, help="rattle the initial positions by this many Angstrom (seeded)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    calc = StudentForceFieldCalculator(ROOT / "tests" / "golden" / "weights_original.npz", device="cuda:0")
    report = {}
    for name in args.systems.split(","):
        atoms = system(name)
        masses = atoms.get_masses()
        for seed in [int(s) for s in args.seeds.split(",")]:
            v0 = md.maxwell_boltzmann(masses, 300.0, np.random.default_rng(seed), atoms.get_positions(), zero_rotation=True)
            x0 = atoms.get_positions() + args.perturb * np.random.default_rng(1000 + seed).normal(size=(len(atoms), 3))
            dev = md.DeviceMD(calc.model, atoms.numbers, x0, v0, masses, 0.5)
            tot = dev.run(args.steps)["total"]
            e0 = abs(tot[0])
            t = np.arange(len(tot))
            slope = np.polyfit(t, tot, 1)[0]
            row = {"endpoint_drift_percent": 100 * (tot[-1] - tot[0]) / e0,
                   "max_abs_deviation_percent": 100 * float(np.max(np.abs(tot - tot[0]))) / e0,
                   "std_percent": 100 * float(np.std(tot)) / e0,
                   "linear_fit_drift_percent": 100 * slope * (len(tot) - 1) / e0}
            report[f"{name}/seed{seed}"] = row
            print(f"{name:8s} seed {seed}: end-point {row['endpoint_drift_percent']:+.4f} %   max |E-E0| {row['max_abs_deviation_percent']:.4f} %   "
                  f"std {row['std_percent']:.4f} %   line fit {row['linear_fit_drift_percent']:+.4f} %", flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
