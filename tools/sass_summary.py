"""Static evidence for the built library (no GPU needed): registers / shared memory / spills per kernel
from `cuobjdump -res-usage`, and the count of the SASS mnemonics that prove which hardware paths the
code uses, from `cuobjdump -sass`.  Usage: python tools/sass_summary.py [libmlffd.so] > profiles/sass_rN.txt"""
import re, subprocess, sys
from collections import Counter, defaultdict
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parents[1] / "mlff_distiller_b200/csrc/libmlffd.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "HMMA", "LDS.128", "LDG.E.128",
         "STG.E.128", "RED", "ATOM", "SHFL", "BAR.SYNC", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage, fn = {}, None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    if fn and "REG:" in line:
        usage[fn] = {k: int(v) for k, v in re.findall(r"(\w+):(\d+)", line)}
        fn = None

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, fn = defaultdict(Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                counts[fn][w] += 1

names = demangle(sorted(set(usage) | set(counts)))


def short(n):
    d = names.get(n, n)
    d = re.sub(r"^void ", "", d)
    d = re.sub(r"\(.*$", "", d)
    return d.replace("mlffd::", "").replace("(anonymous namespace)::", "")


tot = Counter()
for c in counts.values():
    tot.update(c)
print(f"# {lib}")
print("# whole library: " + ", ".join(f"{w} {tot[w]}" for w in WATCH if tot[w]) + f"; instructions {tot['_total']}; kernels {len(usage)}")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 = fma.rn.f32x2; "
      "STL/LDL = local-memory spills")
print(f"{'kernel':90s} {'regs':>4s} {'smem':>6s} {'stack':>5s} {'instr':>6s}  notable")
for n in sorted(usage, key=short):
    if "cub" in short(n) and "Scan" not in short(n):
        pass
    u, c = usage[n], counts.get(n, Counter())
    notable = " ".join(f"{w}:{c[w]}" for w in WATCH if c[w] and w not in ("FFMA", "SHFL", "BAR.SYNC"))
    print(f"{short(n)[:90]:90s} {u.get('REG', 0):4d} {u.get('SHARED', 0):6d} {u.get('STACK', 0):5d} {c['_total']:6d}  {notable}")
