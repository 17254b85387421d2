#!/usr/bin/env python
"""Accuracy (vs the committed FP64 golden outputs of the reference) and per-stage device time of
every arithmetic mode of the dense layers: fp32 | tc | tc_fp16 | tc_bf16.
Usage: python tools/precision_report.py [--variants original,tiny,ultra_tiny] [--out file.json]"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import golden_cases, load_golden  # noqa: E402
from mlff_distiller_b200 import synthetic  # noqa: E402
from mlff_distiller_b200.student_model import StudentForceField  # noqa: E402


def run(model, z, pos, off):
    z_d = torch.from_numpy(np.asarray(z, dtype=np.int32)).cuda()
    p_d = torch.from_numpy(np.asarray(pos, dtype=np.float32)).cuda()
    o_d = torch.from_numpy(np.asarray(off, dtype=np.int32)).cuda()
    e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(off) - 1)
    return e.double().cpu().numpy(), f.double().cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="original,tiny,ultra_tiny")
    ap.add_argument("--modes", default="fp32,tc,tc_fp16,tc_bf16")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    report = {}
    structs = synthetic.druglike_batch(1024)
    zc, pc, oc = synthetic.concatenate(structs)
    for variant in args.variants.split(","):
        gold = load_golden(variant)
        for mode in args.modes.split(","):
            model = StudentForceField.load(ROOT / "tests" / "golden" / f"weights_{variant}.npz", device="cuda:0",
                                           precision=mode)
            worst_e = worst_f = rel_f = 0.0
            for case in golden_cases(gold):
                z, pos, off = gold[f"{case}_numbers"], gold[f"{case}_positions"], gold[f"{case}_offsets"]
                e, f = run(model, z, pos, off)
                worst_e = max(worst_e, float(np.max(np.abs(e - gold[f"{case}_energy64"]) / np.diff(off))))
                if not case.endswith("_exact"):
                    df = float(np.max(np.abs(f - gold[f"{case}_forces64"])))
                    worst_f = max(worst_f, df)
                    fmax = float(np.max(np.abs(gold[f"{case}_forces64"])))
                    if fmax > 0.0:
                        rel_f = max(rel_f, df / fmax)
            # device time of one C2 step
            z_d = torch.from_numpy(zc.astype(np.int32)).cuda()
            p_d = torch.from_numpy(pc.astype(np.float32)).cuda()
            o_d = torch.from_numpy(oc.astype(np.int32)).cuda()
            e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(structs), max_atoms=50)
            eng = model.engine()
            for _ in range(3):
                eng.energy_forces_async(z_d, p_d, o_d, len(structs), e, f)
            torch.cuda.synchronize()
            eng.profile_enable(True)
            for _ in range(args.steps):
                eng.energy_forces_async(z_d, p_d, o_d, len(structs), e, f)
            prof = eng.profile_read()
            eng.profile_enable(False)
            stages = {k: round(v["ms"] / args.steps, 4) for k, v in prof["stages"].items() if v["launches"]}
            row = {"max_dE_per_atom_eV": worst_e, "max_dF_eV_per_A": worst_f, "max_dF_rel_to_max_F": rel_f,
                   "c2_ms_per_step": round(sum(stages.values()), 4), "c2_stages_ms": stages}
            report[f"{variant}/{mode}"] = row
            print(f"{variant:11s} {mode:8s} dE/atom {worst_e:.2e}  dF {worst_f:.2e} (rel {rel_f:.1e})  "
                  f"C2 {row['c2_ms_per_step']:.3f} ms  filter {stages.get('filter', 0):.3f}  "
                  f"upd {stages.get('update_fwd', 0) + stages.get('update_bwd', 0):.3f}", flush=True)
            del model
    if args.out:
        Path(args.out).write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
