#!/usr/bin/env python
"""Launch the filter-table kernel alone on C2-like distances (for ncu).  --precision fp32|tc"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mlff_distiller_b200.student_model import StudentForceField  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tc"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 513281
model = StudentForceField.load(ROOT / "tests" / "golden" / "weights_original.npz", device="cuda:0", precision=prec)
eng = model.engine()
d = torch.from_numpy(np.random.default_rng(0).uniform(0.9, 5.0, P).astype(np.float32)).cuda()
for rep in range(3):
    for layer in (1, 2):
        f, df = eng.filter_table(layer, d)
torch.cuda.synchronize()
start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
for rep in range(10):
    f, df = eng.filter_table(1, d)
end.record()
torch.cuda.synchronize()
print(f"{prec}: {start.elapsed_time(end) / 10:.3f} ms per launch, P={P}")
