#!/usr/bin/env python
"""C5 through ONE process: StudentForceFieldCalculator(device_ids=all GPUs).evaluate_arrays on the whole
list (one context + streams + pinned staging per GPU, driven by host threads)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator  # noqa: E402

cache, total = sys.argv[1], int(sys.argv[2])
with np.load(cache) as z:
    counts = z["counts"][:total].astype(np.int64)
    n = int(counts.sum())
    numbers, pos = z["numbers"][:n].astype(np.int64), z["positions"][:n].astype(np.float64)
out = {"config": "C5 ragged sweep, one process", "structures": total, "atoms": n, "results": {}}
for ids in ([0], list(range(torch.cuda.device_count()))):
    for variant in ("original", "tiny", "ultra_tiny"):
        calc = StudentForceFieldCalculator(ROOT / "tests" / "golden" / f"weights_{variant}.npz", device_ids=ids)
        warm = int(np.searchsorted(np.cumsum(counts), 200000))
        calc.evaluate_arrays(numbers[: counts[:warm].sum()], pos[: counts[:warm].sum()], counts[:warm])
        best = 1e9
        for _ in range(3):
            for d in ids:
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            e, f = calc.evaluate_arrays(numbers, pos, counts)
            best = min(best, time.perf_counter() - t0)
        out["results"][f"{variant}_{len(ids)}gpu"] = {"structures_per_s": total / best, "seconds": best,
                                                      "energy_checksum": float(np.asarray(e, np.float64).sum())}
        del calc
print(json.dumps(out))
