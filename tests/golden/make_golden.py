#!/usr/bin/env python
"""Generate the golden fixtures by running the REAL reference (container only).

Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference (read-only).  Writes, for each variant (original, tiny, ultra_tiny):

  tests/golden/weights_<variant>.npz   trained FP32 tensors read from the reference's ONNX
                                       exports (models/original_model.onnx,
                                       benchmarks/{tiny,ultra_tiny}_model.onnx) + __config__
  tests/golden/golden_<variant>.npz    inputs and the reference's own outputs:
      <case>_numbers / _positions / _offsets    inputs (positions float32)
      <case>_energy32 / _forces32              reference FP32 (predict_energy_and_forces for a
                                               single structure; forward(batch)+autograd for
                                               batches, as inference/ase_calculator.py:757-763)
      <case>_energy64 / _forces64              same modules cast to FP64 (ground truth)
      <case>_edge_index                         radius_graph_native output

The reference model is imported by file path (oracle/reference_loader.py); nothing from it is
copied into the repository.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from mlff_distiller_b200 import checkpoint, synthetic  # noqa: E402
from oracle import reference_loader  # noqa: E402

REF = reference_loader.REFERENCE_ROOT
VARIANTS = {
    "original": REF / "models/original_model.onnx",
    "tiny": REF / "benchmarks/tiny_model.onnx",
    "ultra_tiny": REF / "benchmarks/ultra_tiny_model.onnx",
}


def cases():
    out = {}
    w = synthetic.water()
    out["h2o"] = [w]
    out["benzene"] = [synthetic.benzene()]
    # reference unit-test fixture water (tests/unit/test_student_model.py:71-93)
    out["fixture_water"] = [synthetic.Structure([8, 1, 1], [[0, 0, 0], [0.96, 0, 0], [-0.24, 0.93, 0]])]
    out["single_atom"] = [synthetic.Structure([6], [[0.1, 0.2, 0.3]])]
    # one atom beyond the cutoff of everything else: zero-norm vector features in update
    iso = synthetic.druglike(77, 13)
    pos = np.vstack([iso.positions, [[30.0, 0.0, 0.0]]])
    out["isolated"] = [synthetic.Structure(np.append(iso.numbers, 8), pos)]
    out["drug50"] = [synthetic.druglike(1000, 50)]
    out["batch4x50"] = synthetic.druglike_batch(4)
    out["ragged"] = synthetic.druglike_batch(5, first=10, ragged=True) + [synthetic.water()]
    # The exact chain is mirror-symmetric: vector features cancel to rounding noise and the
    # gradient of their norm is ill-conditioned (reference FP32 vs FP64 forces differ by >1 eV/A).
    # Golden cases therefore use a seeded 0.05 A rattle; "chain300_exact" pins the energy only.
    for key, units, seed in (("chain30", 10, 7), ("chain300", 100, 8)):
        c = synthetic.alkane_chain(units)
        c.positions = c.positions + np.random.default_rng(seed).normal(0.0, 0.05, c.positions.shape)
        out[key] = [c]
    out["chain300_exact"] = [synthetic.alkane_chain(100)]
    return out


def run_reference(model, z, pos, offsets, dtype):
    model = model.to(dtype)
    zt = torch.from_numpy(z)
    pt = torch.from_numpy(pos).to(dtype)
    nb = len(offsets) - 1
    if nb == 1:
        e, f = model.predict_energy_and_forces(zt, pt.clone())
        e = e.detach().reshape(1)
    else:
        batch = torch.from_numpy(np.repeat(np.arange(nb), np.diff(offsets)))
        p = pt.clone().requires_grad_(True)
        e = model(zt, p, cell=None, pbc=None, batch=batch)
        f = -torch.autograd.grad(e, p, grad_outputs=torch.ones_like(e))[0]
        e = e.detach()
    return e.numpy(), f.detach().numpy()


def main():
    mod = reference_loader.load_reference_module("student_model")
    out_dir = ROOT / "tests" / "golden"
    for name, onnx_path in VARIANTS.items():
        state, cfg, _ = checkpoint.load_any(onnx_path)
        np.savez(out_dir / f"weights_{name}.npz", __config__=json.dumps(cfg.as_dict()), **state)
        gold = {}
        for case, structs in cases().items():
            z, pos64, offsets = synthetic.concatenate(structs)
            pos32 = pos64.astype(np.float32)
            gold[f"{case}_numbers"] = z
            gold[f"{case}_positions"] = pos32
            gold[f"{case}_offsets"] = offsets
            model = reference_loader.build_reference_model(state, cfg)
            e32, f32 = run_reference(model, z, pos32, offsets, torch.float32)
            model = reference_loader.build_reference_model(state, cfg)
            e64, f64 = run_reference(model, z, pos32, offsets, torch.float64)
            batch = torch.from_numpy(np.repeat(np.arange(len(offsets) - 1), np.diff(offsets)))
            ei = mod.radius_graph_native(torch.from_numpy(pos32), cfg.cutoff, batch)
            gold[f"{case}_energy32"], gold[f"{case}_forces32"] = e32, f32
            gold[f"{case}_energy64"], gold[f"{case}_forces64"] = e64, f64
            gold[f"{case}_edge_index"] = ei.numpy().astype(np.int32)
            print(f"{name:11s} {case:14s} N={len(z):4d} E={ei.shape[1]:6d} "
                  f"E32={e32.sum():.6f} max|F32-F64|={np.abs(f32 - f64).max():.2e}")
        np.savez_compressed(out_dir / f"golden_{name}.npz", **gold)


if __name__ == "__main__":
    main()
