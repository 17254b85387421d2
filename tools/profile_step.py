#!/usr/bin/env python
"""Run a few device-resident E+F steps of a named workload (for ncu / sanitizer captures).
Usage: python tools/profile_step.py [--workload c2|c3|c4] [--variant original] [--steps 4]"""
import argparse
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
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--variant", default="original")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default="tc")
    args = ap.parse_args()
    model = StudentForceField.load(ROOT / "tests" / "golden" / f"weights_{args.variant}.npz", device="cuda:0",
                                   pbc_mode="minimum_image", precision=args.precision)
    cells = pbc = None
    if args.workload == "c2":
        structs = synthetic.druglike_batch(args.batch)
    elif args.workload == "c1":
        structs = [synthetic.water()]
    elif args.workload == "c3":
        structs = [synthetic.alkane_chain(100)]
    else:
        structs = [synthetic.water_box()]
        cells, pbc = model.pack_cells(torch.from_numpy(structs[0].cell), torch.from_numpy(structs[0].pbc), 1, "cuda:0")
    z, pos, off = synthetic.concatenate(structs)
    z_d = torch.from_numpy(z.astype(np.int32)).cuda()
    p_d = torch.from_numpy(pos.astype(np.float32)).cuda()
    o_d = torch.from_numpy(off.astype(np.int32)).cuda()
    e, f = model.energy_and_forces_packed(z_d, p_d, o_d, len(structs), cells, pbc)
    st = model.engine().status()
    print(f"N={len(z)} E={st.num_edges} P={st.num_pairs} maxdeg={st.max_degree} E0={float(e[0]):.4f}")
    eng = model.engine()
    eng.profile_enable(True)
    for _ in range(args.steps):
        eng.energy_forces_async(z_d, p_d, o_d, len(structs), e, f, cells, pbc)
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile_enable(False)
    tot = 0.0
    for k, v in prof["stages"].items():
        if v["launches"]:
            print(f"  {k:12s} {1e3 * v['ms'] / args.steps:9.1f} us/step  ({v['launches'] // args.steps} launches)")
            tot += 1e3 * v["ms"] / args.steps
    print(f"  total        {tot:9.1f} us/step (device time between events)")
    print("done")


if __name__ == "__main__":
    main()
