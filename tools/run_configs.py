#!/usr/bin/env python
"""Run the non-headline BASELINE.json configs on one B200 and write a JSON report.

  C1  H2O, 1000-step velocity-Verlet NVE (dt 0.5 fs, T0 300 K, seed 42): steps/s, ns/day, drift %
  C3  300-atom CH2 chain (the reference's peptide generator extended, rattled 0.02 A),
      NVE 20 000 steps @ 0.5 fs (10 ps): steps/s, ns/day, drift %
  C4  9 999-atom periodic water box, neighbour list rebuilt every step: ms/step split
  C5  ragged screening sweep (n ~ U{20..80}) for Original / Tiny / Ultra-tiny: structures/s

Every MD loop calls the calculator once per step with fresh positions (no ASE cache hits).
The CPU column times the oracle port of the reference path on the host for a bounded number of
steps (reported baseline, not a target).  Usage: python tools/run_configs.py [--out path] [--quick]
"""
import argparse
import json
import os
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


def load_state(variant):
    with np.load(W / f"weights_{variant}.npz") as z:
        return {k: z[k] for k in z.files if not k.startswith("__")}, json.loads(str(z["__config__"]))


def nve(calc, atoms, steps, temperature, seed, dt_fs=0.5):
    rng = np.random.default_rng(seed)
    masses = atoms.get_masses()
    v0 = md.maxwell_boltzmann(masses, temperature, rng, atoms.get_positions(), zero_rotation=True)
    work = atoms.copy()

    def force_fn(x):
        work.set_positions(x)
        calc.calculate(work, ["energy", "forces"])
        return calc.results["energy"], calc.results["forces"]

    for _ in range(3):   # warm-up: workspace sizing (eager calls), graph capture on the fourth call, replays
        force_fn(atoms.get_positions() + 1e-6)
        force_fn(atoms.get_positions())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = md.velocity_verlet(force_fn, atoms.get_positions(), v0, masses, steps, dt_fs)
    dt = time.perf_counter() - t0
    sps = steps / dt
    return {"steps": steps, "dt_fs": dt_fs, "T0_K": temperature, "steps_per_s": sps,
            "us_per_step": 1e6 / sps, "ns_per_day": md.ns_per_day(sps, dt_fs),
            "drift_percent": out["drift_percent"], "E_total_first": float(out["total"][0]),
            "E_total_last": float(out["total"][-1]),
            "max_abs_dE_total": float(np.abs(out["total"] - out["total"][0]).max())}


def nve_device(calc, atoms, steps, temperature, seed, dt_fs=0.5, cell=None, pbc=None):
    """Same trajectory with the integrator on the GPU (CUDA-graph replay, no per-step copies)."""
    rng = np.random.default_rng(seed)
    masses = atoms.get_masses()
    v0 = md.maxwell_boltzmann(masses, temperature, rng, atoms.get_positions(), zero_rotation=True)
    dev = md.DeviceMD(calc.model, atoms.numbers, atoms.get_positions(), v0, masses, dt_fs, cell=cell, pbc=pbc)
    dev.run(20)  # graph capture + warm-up (these steps are part of the trajectory)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = dev.run(steps - 20)
    dt = time.perf_counter() - t0
    sps = (steps - 20) / dt
    return {"steps": steps, "steps_per_s": sps, "us_per_step": 1e6 / sps, "ns_per_day": md.ns_per_day(sps, dt_fs),
            "drift_percent": out["drift_percent"], "E_total_first": float(out["total"][0]),
            "E_total_last": float(out["total"][-1])}


def cpu_steps(variant, atoms, steps, cells=None, pbc=None):
    state, cfg = load_state(variant)
    torch.set_num_threads(os.cpu_count() or 1)
    z, pos = atoms.numbers, atoms.positions.astype(np.float32)
    po.evaluate(state, cfg["cutoff"], z, pos, [0, len(z)], cells, pbc)
    t0 = time.perf_counter()
    for i in range(steps):
        po.evaluate(state, cfg["cutoff"], z, pos + np.float32(1e-4 * i), [0, len(z)], cells, pbc)
    dt = time.perf_counter() - t0
    return {"steps": steps, "steps_per_s": steps / dt, "ns_per_day": md.ns_per_day(steps / dt),
            "cores": torch.get_num_threads(), "kind": "port (oracle/painn_oracle.py)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "configs.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--precision", default="tc")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    report = {"gpu": torch.cuda.get_device_name(0), "precision": args.precision}

    def want(name):
        return only is None or name in only

    if want("c1"):
        calc = StudentForceFieldCalculator(W / "weights_original.npz", device="cuda:0", precision=args.precision)
        r = nve(calc, synthetic.water(), 200 if args.quick else 1000, 300.0, 42)
        r["device_md"] = nve_device(calc, synthetic.water(), 200 if args.quick else 1000, 300.0, 42)
        r["device_md_10ps"] = nve_device(calc, synthetic.water(), 2000 if args.quick else 20000, 300.0, 42)
        r["benzene_device_md_10ps"] = nve_device(calc, synthetic.benzene(), 2000 if args.quick else 20000, 300.0, 42)
        r["cpu_reference"] = cpu_steps("original", synthetic.water(), 50 if args.quick else 200)
        report["C1_h2o_nve"] = r
        print("C1", json.dumps(r), flush=True)

    if want("c3"):
        calc = StudentForceFieldCalculator(W / "weights_original.npz", device="cuda:0", precision=args.precision)
        chain = synthetic.alkane_chain(100)
        chain.positions = chain.positions + np.random.default_rng(8).normal(0.0, 0.02, chain.positions.shape)
        r = nve(calc, chain, 2000 if args.quick else 20000, 300.0, 42)
        r["device_md"] = nve_device(calc, chain, 2000 if args.quick else 20000, 300.0, 42)
        r["cpu_reference"] = cpu_steps("original", chain, 5 if args.quick else 30)
        report["C3_chain300_nve"] = r
        print("C3", json.dumps(r), flush=True)

    if want("c4"):
        calc = StudentForceFieldCalculator(W / "weights_original.npz", device="cuda:0", precision=args.precision,
                                           pbc_mode="minimum_image")
        box = synthetic.water_box()
        work = box.copy()
        calc.calculate(work)
        eng = calc.model.engine()
        st = eng.status()
        rng = np.random.default_rng(0)
        steps = 5 if args.quick else 20
        eng.profile_enable(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            work.set_positions(box.positions + rng.normal(0, 0.01, box.positions.shape))
            calc.calculate(work)
        dt = time.perf_counter() - t0
        prof = eng.profile_read()
        eng.profile_enable(False)
        stages = {k: v["ms"] / steps for k, v in prof["stages"].items() if v["launches"]}
        fwd = sum(stages.get(k, 0) for k in ("embedding", "filter", "message_fwd", "update_fwd", "readout", "energy_sum"))
        rev = sum(stages.get(k, 0) for k in ("update_bwd", "message_bwd", "force"))
        r = {"atoms": len(box), "edges": int(st.num_edges), "ms_per_step_wall": 1e3 * dt / steps,
             "ms_neighbor": stages.get("neighbor"), "ms_forward": fwd, "ms_reverse": rev, "stages_ms": stages,
             "steps_per_s": steps / dt, "ns_per_day": md.ns_per_day(steps / dt),
             "energy_eV": calc.results["energy"], "max_force": float(np.abs(calc.results["forces"]).max())}
        r["device_md"] = nve_device(calc, box, 120 if args.quick else 520, 300.0, 42, cell=box.cell, pbc=box.pbc)
        if not args.quick:
            r["cpu_reference"] = cpu_steps("original", box, 1, box.cell[None], box.pbc[None])
        report["C4_water_box_10k"] = r
        print("C4", json.dumps(r), flush=True)

    if want("c5"):
        n_struct = 2048 if args.quick else 20480
        structs = synthetic.druglike_batch(n_struct, first=0, ragged=True)
        numbers, pos, offsets = synthetic.concatenate(structs)
        counts = np.diff(offsets)
        res = {}
        for variant in ("original", "tiny", "ultra_tiny"):
            calc = StudentForceFieldCalculator(W / f"weights_{variant}.npz", device="cuda:0", precision=args.precision)
            chunk = 2048

            def chunks():
                for s0 in range(0, n_struct, chunk):
                    s1 = min(n_struct, s0 + chunk)
                    yield numbers[offsets[s0]:offsets[s1]], pos[offsets[s0]:offsets[s1]], counts[s0:s1]

            for _ in calc.evaluate_stream(iter(list(chunks())[:2])):   # size workspace + staging before timing
                pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e_all = [e for e, f in calc.evaluate_stream(chunks())]   # host arrays in, host arrays out, two chunks in flight
            dt = time.perf_counter() - t0
            res[variant] = {"structures": n_struct, "atoms": int(offsets[-1]), "structures_per_s": n_struct / dt,
                            "atoms_per_s": int(offsets[-1]) / dt, "mean_energy_per_atom": float(np.concatenate(e_all).sum() / offsets[-1])}
        report["C5_sweep_1gpu"] = res
        print("C5", json.dumps(res), flush=True)

    Path(args.out).parent.mkdir(exist_ok=True, parents=True)
    Path(args.out).write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
