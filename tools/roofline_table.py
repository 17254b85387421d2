"""Per-kernel roofline table from an ncu summary written by tools/ncu_summary.py (profiles/ncu_full_*.txt):
DRAM bytes / duration against the measured HBM peak, tensor-pipe and L1-data-pipe utilisation as ncu reports
them.  ncu's per-launch durations are cold-cache and serialised (B200_PROFILING.md): the fractions say which
pipe a kernel leans on; the step-level numbers are bench.py's.
Usage: python tools/roofline_table.py profiles/ncu_full_r2_final.txt > profiles/roofline_r2_final.md"""
import json
import re
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
src = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "profiles" / "ncu_full_r2_final.txt")
peaks = ROOT / "MEASURED_PEAKS.json"
hbm = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
launches, cur = [], None
for line in src.read_text().splitlines():
    m = re.match(r"\s+kernel:\s+(?:void\s+)?(.+)", line)
    if m:
        name = re.sub(r"\(.*", "", m.group(1)).strip()
        cur = {"kernel": name}
        launches.append(cur)
        continue
    m = re.match(r"\s+(\S+)\s+([0-9.eE+-]+)\s*(\S*)", line)
    if m and cur is not None:
        cur[m.group(1)] = float(m.group(2)) * UNITS.get(m.group(3), 1.0)

groups = OrderedDict()
for l in launches:
    groups.setdefault(l["kernel"], []).append(l)


def mean(rows, key):
    vals = [r[key] for r in rows if key in r]
    return sum(vals) / len(vals) if vals else float("nan")


print(f"# Per-kernel roofline fractions from `{src.name}` (ncu --set full, one C2 step; HBM peak {hbm:.0f} GB/s measured)\n")
print("| kernel | launches | mean µs | DRAM MB / launch | DRAM GB/s | frac of HBM peak | tensor pipe % | FMA pipe % | L1 data-pipe wavefronts % | warps active % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, rows in groups.items():
    t = mean(rows, "gpu__time_duration.sum")
    b = mean(rows, "dram__bytes_read.sum") + mean(rows, "dram__bytes_write.sum")
    gbs = b / t / 1e9 if t > 0 else float("nan")
    print(f"| `{name}` | {len(rows)} | {t * 1e6:.1f} | {b / 1e6:.1f} | {gbs:.0f} | {gbs / hbm:.2f} | "
          f"{mean(rows, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{mean(rows, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{mean(rows, 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{mean(rows, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{int(mean(rows, 'launch__registers_per_thread'))} |")
