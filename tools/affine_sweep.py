#!/usr/bin/env python
"""Sweep the row order of the pipelined message kernels on C2 (structure-affine block sizes /
units per structure vs grid-stride): per-stage device time and bit-identity of the results.
Usage: python tools/affine_sweep.py [--configs "0:0:1,24:16:1,..."]   (fwd warps : bwd warps : parts)"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mlff_distiller_b200 import synthetic  # noqa: E402
from mlff_distiller_b200.student_model import StudentForceField  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="0:0:1,24:16:1,12:16:1,8:8:1,8:8:2,24:16:2,12:8:2")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--ragged", action="store_true")
    args = ap.parse_args()
    structs = synthetic.druglike_batch(1024, ragged=args.ragged)
    z, pos, off = synthetic.concatenate(structs)
    z_d = torch.from_numpy(z.astype(np.int32)).cuda()
    p_d = torch.from_numpy(pos.astype(np.float32)).cuda()
    o_d = torch.from_numpy(off.astype(np.int32)).cuda()
    base = None
    for cfg in args.configs.split(","):
        fw, bw, parts = cfg.split(":")
        os.environ["MLFFD_AFFINE_FWD"], os.environ["MLFFD_AFFINE_BWD"], os.environ["MLFFD_AFFINE_PARTS"] = fw, bw, parts
        model = StudentForceField.load(ROOT / "tests" / "golden" / "weights_original.npz", device="cuda:0")
        e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(structs), max_atoms=int(np.diff(off).max()))
        torch.cuda.synchronize()
        if base is None:
            base = (e.clone(), f.clone())
        same = bool(torch.equal(e, base[0]) and torch.equal(f, base[1]))
        eng = model.engine()
        for _ in range(3):
            eng.energy_forces_async(z_d, p_d, o_d, len(structs), e, f)
        torch.cuda.synchronize()
        eng.profile_enable(True)
        for _ in range(args.steps):
            eng.energy_forces_async(z_d, p_d, o_d, len(structs), e, f)
        prof = eng.profile_read()
        eng.profile_enable(False)
        st = {k: v["ms"] / args.steps for k, v in prof["stages"].items() if v["launches"]}
        print(f"fwd W={fw:>2s} bwd W={bw:>2s} parts={parts}: msg_fwd {st['message_fwd']:.3f}  msg_bwd {st['message_bwd']:.3f}  "
              f"step {sum(st.values()):.3f} ms  bit-identical to first config: {same}", flush=True)
        del model, eng


if __name__ == "__main__":
    main()
