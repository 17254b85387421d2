#!/usr/bin/env python
"""Quick GPU sanity run: every variant on a few golden cases, errors vs the reference outputs.
Usage: python tools/sanity_gpu.py [variant ...]   (also handy under compute-sanitizer)"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import load_golden, load_weights  # noqa: E402
from mlff_distiller_b200.checkpoint import infer_config  # noqa: E402
from mlff_distiller_b200.student_model import StudentForceField  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    precision = "tc" if "--tc" in sys.argv else "fp32"
    variants = args or ["original", "tiny", "ultra_tiny"]
    for variant in variants:
        state, cfg = load_weights(variant)
        gold = load_golden(variant)
        model = StudentForceField.from_state(state, infer_config(state, cfg), "cuda:0", precision=precision)
        for case in ("h2o", "single_atom", "isolated", "drug50", "ragged", "chain300"):
            z = torch.from_numpy(gold[f"{case}_numbers"].astype(np.int32)).cuda()
            pos = torch.from_numpy(gold[f"{case}_positions"]).cuda()
            off = torch.from_numpy(gold[f"{case}_offsets"].astype(np.int32)).cuda()
            nb = len(gold[f"{case}_offsets"]) - 1
            e, f = model.energy_and_forces_packed(z, pos, off, nb)
            st = model.engine().status()
            e = e.double().cpu().numpy()
            f = f.double().cpu().numpy()
            de = np.max(np.abs(e - gold[f"{case}_energy64"]) / np.diff(gold[f"{case}_offsets"]))
            df = np.max(np.abs(f - gold[f"{case}_forces64"]))
            print(f"{variant:10s} {case:12s} E={st.num_edges:6d} (ref {gold[f'{case}_edge_index'].shape[1]:6d}) "
                  f"dE/atom={de:.3e} dF={df:.3e} |F|max={np.abs(f).max():.3f} e0={e[0]:.6f} ref={gold[f'{case}_energy32'][0]:.6f}",
                  flush=True)
    torch.cuda.synchronize()
    print("sanity done")


if __name__ == "__main__":
    main()
