import sys, time, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from mlff_distiller_b200 import md, synthetic
from mlff_distiller_b200.ase_calculator import StudentForceFieldCalculator
W = ROOT / "tests" / "golden"
for fm in ("spline", "table"):
    calc = StudentForceFieldCalculator(W / "weights_original.npz", device="cuda:0", filter_mode=fm)
    for name, atoms in (("h2o", synthetic.water()), ("benzene", synthetic.benzene()), ("drug50", synthetic.druglike(1000, 50)), ("chain300", synthetic.alkane_chain(100))):
        if name == "chain300":
            atoms.positions = atoms.positions + np.random.default_rng(8).normal(0.0, 0.02, atoms.positions.shape)
        m = atoms.get_masses()
        v0 = md.maxwell_boltzmann(m, 300.0, np.random.default_rng(42), atoms.get_positions(), zero_rotation=True)
        sim = md.DeviceMD(calc.model, atoms.numbers, atoms.get_positions(), v0, m, 0.5)
        e0 = int(sim.eng.status().num_edges)
        sim.run(50)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = sim.run(1000)
        dt = time.perf_counter() - t0
        print(fm, name, "edges t0", e0, "edges end", int(sim.eng.status().num_edges), "us/step %.1f" % (dt / 1000 * 1e6), "drift %.4f" % out["drift_percent"], flush=True)
