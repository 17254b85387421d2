#!/usr/bin/env python
"""Generate the C5 structure list once (all host cores) so that several sweep runs can share it:
    python tools/make_sweep_cache.py 100000 /tmp/sweep.npz"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

if __name__ == "__main__":
    total, out = int(sys.argv[1]), sys.argv[2]
    numbers, pos, counts = bench.sweep_shard(0, total, os.cpu_count() or 1)
    np.savez(out, numbers=numbers.astype(np.int16), positions=pos.astype(np.float32), counts=counts.astype(np.int32))
    print(f"{total} structures, {len(numbers)} atoms -> {out}")
