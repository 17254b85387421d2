"""compute-sanitizer target: one small pass through the throughput kernels (small-system path off), the skin
list, the virial and the small-system kernels.  Usage (GPU box):
    compute-sanitizer --tool memcheck python tools/sanitize_gpu.py
    compute-sanitizer --tool racecheck python tools/sanitize_gpu.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mlff_distiller_b200 import synthetic                                   # noqa: E402
from mlff_distiller_b200.student_model import StudentForceField             # noqa: E402
from mlff_distiller_b200.checkpoint import infer_config                     # noqa: E402


def load_weights(variant):
    import json
    with np.load(Path(__file__).resolve().parents[1] / "tests" / "golden" / f"weights_{variant}.npz") as z:
        state = {k: z[k] for k in z.files if not k.startswith("__")}
        cfg = json.loads(str(z["__config__"]))
    return state, cfg


def run(variant, env, **kw):
    os.environ.update(env)
    try:
        state, cfg = load_weights(variant)
        m = StudentForceField.from_state(state, infer_config(state, cfg), "cuda:0", **kw)
        structs = synthetic.druglike_batch(6, first=40, ragged=True) + [synthetic.water()]
        z, pos, off = synthetic.concatenate(structs)
        dev = "cuda:0"
        zd = torch.from_numpy(np.asarray(z, dtype=np.int32)).to(dev)
        pd = torch.from_numpy(pos.astype(np.float32)).to(dev)
        od = torch.from_numpy(np.asarray(off, dtype=np.int32)).to(dev)
        e, f = m.energy_and_forces_packed(zd, pd, od, len(off) - 1, None, None)
        w = m.virial_of_last_call(od, len(off) - 1)
        torch.cuda.synchronize()
        assert torch.isfinite(e).all() and torch.isfinite(f).all() and torch.isfinite(w).all()
        box = synthetic.water_box(n_mol=96, seed=7)
        cd, bd = StudentForceField.pack_cells(torch.from_numpy(box.cell[None]), torch.from_numpy(box.pbc[None]), 1, dev)
        mp = StudentForceField.from_state(state, infer_config(state, cfg), dev, pbc_mode="minimum_image", **kw)
        zb = torch.from_numpy(np.asarray(box.numbers, dtype=np.int32)).to(dev)
        pb = torch.from_numpy(box.positions.astype(np.float32)).to(dev)
        ob = torch.tensor([0, len(box.numbers)], dtype=torch.int32, device=dev)
        for step in range(3):
            eb, fb = mp.energy_and_forces_packed(zb, pb + 0.05 * step, ob, 1, cd, bd)
        torch.cuda.synchronize()
        assert torch.isfinite(eb).all() and torch.isfinite(fb).all()
        print("ok", variant, env, kw, float(e.sum()), float(eb[0]))
    finally:
        for k in env:
            os.environ.pop(k, None)


if __name__ == "__main__":
    run("original", {"MLFFD_SMALL_ROWS": "0"})          # throughput kernels (spline rows, tcgen05 update block)
    run("original", {})                                   # small-system kernels
    run("ultra_tiny", {"MLFFD_SMALL_ROWS": "0"}, skin=1.0)
