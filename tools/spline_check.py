#!/usr/bin/env python
"""GPU check of the spline filter mode: spline vs the exact filter stage, golden E/F parity per variant,
and agreement with the table mode.  Writes gpurun_out/spline_check.json."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import golden_cases, load_golden, load_weights  # noqa: E402
from mlff_distiller_b200.checkpoint import infer_config  # noqa: E402
from mlff_distiller_b200.student_model import StudentForceField  # noqa: E402


def run(model, z, pos, off):
    dev = "cuda:0"
    z_d = torch.from_numpy(np.asarray(z, dtype=np.int32)).to(dev)
    p_d = torch.from_numpy(np.asarray(pos, dtype=np.float32)).to(dev)
    o_d = torch.from_numpy(np.asarray(off, dtype=np.int32)).to(dev)
    e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(off) - 1)
    return e.double().cpu().numpy(), f.double().cpu().numpy()


report = {}
for variant in ("original", "tiny", "ultra_tiny"):
    state, cfg = load_weights(variant)
    c = infer_config(state, cfg)
    models = {m: StudentForceField.from_state(state, c, "cuda:0", precision="tc", filter_mode=m) for m in ("spline", "table")}
    rc = cfg["cutoff"]
    d = torch.cat([torch.linspace(0.3, rc, 20001), torch.tensor([rc, rc - 1e-6, 0.9572])]).float()
    rep = {"filter": [], "cases": {}}
    eng32 = StudentForceField.from_state(state, c, "cuda:0", precision="fp32", filter_mode="table").engine()
    for l in range(cfg["num_interactions"]):
        fs, dfs = models["spline"].engine().filter_spline(l, d)
        ft, dft = eng32.filter_table(l, d)
        rep["filter"].append({"val_err": float((fs - ft).abs().max()), "der_err": float((dfs - dft).abs().max()),
                              "amp": float(ft.abs().max()), "damp": float(dft.abs().max())})
    gold = load_golden(variant)
    for case in golden_cases(gold):
        z, pos, off = gold[f"{case}_numbers"], gold[f"{case}_positions"], gold[f"{case}_offsets"]
        out = {}
        for m, model in models.items():
            e, f = run(model, z, pos, off)
            n = np.diff(off)
            out[m] = {"dE64_per_atom": float(np.max(np.abs(e - gold[f"{case}_energy64"]) / n)),
                      "dF64": float(np.max(np.abs(f - gold[f"{case}_forces64"])))}
            out[m + "_ef"] = (e, f)
        es, fs_ = out.pop("spline_ef"); et, ft_ = out.pop("table_ef")
        out["spline_vs_table"] = {"dE_per_atom": float(np.max(np.abs(es - et) / np.diff(off))), "dF": float(np.max(np.abs(fs_ - ft_)))}
        out["ref_fp32_vs_fp64_F"] = float(np.max(np.abs(gold[f"{case}_forces32"] - gold[f"{case}_forces64"])))
        rep["cases"][case] = out
    report[variant] = rep
    print(variant, json.dumps(rep, indent=1))
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
(out / "spline_check.json").write_text(json.dumps(report, indent=1))
