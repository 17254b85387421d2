#!/usr/bin/env python
"""Single-trajectory latency of the on-device MD step (CUDA-graph replay) for small systems:
us/step and ns/day at 0.5 fs.  Usage: python tools/md_latency.py [--systems h2o,benzene,drug50,chain300] [--steps 2000]"""
import argparse
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


def system(name):
    if name == "h2o":
        return synthetic.water()
    if name == "benzene":
        return synthetic.benzene()
    if name == "drug50":
        return synthetic.druglike_batch(1, first=7)[0]
    if name == "chain300":
        chain = synthetic.alkane_chain(100)
        chain.positions = chain.positions + np.random.default_rng(8).normal(0.0, 0.02, chain.positions.shape)
        return chain
    raise SystemExit(f"unknown system {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--systems", default="h2o,benzene,drug50,chain300")
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--variant", default="original")
    ap.add_argument("--precision", default="tc")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    calc = StudentForceFieldCalculator(ROOT / "tests" / "golden" / f"weights_{args.variant}.npz", device="cuda:0",
                                       precision=args.precision)
    report = {}
    for name in args.systems.split(","):
        atoms = system(name)
        masses = atoms.get_masses()
        v0 = md.maxwell_boltzmann(masses, 300.0, np.random.default_rng(42), atoms.get_positions(), zero_rotation=True)
        dev = md.DeviceMD(calc.model, atoms.numbers, atoms.get_positions(), v0, masses, 0.5)
        dev.run(20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = dev.run(args.steps)
        dt = time.perf_counter() - t0
        sps = args.steps / dt
        report[name] = {"atoms": len(atoms), "us_per_step": 1e6 / sps, "ns_per_day": md.ns_per_day(sps, 0.5),
                        "drift_percent": out["drift_percent"]}
        print(f"{name:9s} {len(atoms):4d} atoms  {1e6 / sps:7.1f} us/step  {md.ns_per_day(sps, 0.5):7.1f} ns/day  "
              f"drift {out['drift_percent']:+.4f} %", flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
