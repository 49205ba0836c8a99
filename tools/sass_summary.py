"""Per-kernel SASS evidence: counts of the instructions that show what each kernel is built from.

  python tools/sass_summary.py [path/to/libplbm_b200.so] > profiles/rNN_sass_summary.txt

Reads `cuobjdump -sass` of the built library (no GPU needed).  Columns: UBLKCP = cp.async.bulk (bulk async copies),
UTMALDG = TMA tensor loads (cp.async.bulk.tensor), SYNCS = mbarrier operations, LDG.128 / STG.128 = 128-bit global
loads / stores, LDS / STS = shared-memory traffic, DFMA / DADD / DMUL = fp64 pipe, FFMA2 = packed fp32 pairs,
FADD+FMUL+FFMA = scalar fp32, SHFL = warp shuffles, BAR = block barriers, tensor = any HMMA/UTCMMA/… (expected 0)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "periodic_lbm_b200", "libplbm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
demangled = dict(zip(re.findall(r"Function : (\S+)", sass), names))

COLS = ["UBLKCP", "UTMALDG", "SYNCS", "LDG.128", "STG.128", "LDG", "STG", "LDS", "STS", "DFMA", "DADD", "DMUL", "FFMA2", "F32", "SHFL", "BAR", "tensor"]
counts = collections.OrderedDict()
cur = None
unit = ""
arch = set()
for line in sass.split("\n"):
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"identifier = (\S+)", line)
    if m:
        unit = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = counts.setdefault((m.group(1), unit), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    if base in ("UBLKCP", "UTMALDG", "SYNCS", "LDS", "STS", "DFMA", "DADD", "DMUL", "SHFL", "BAR"):
        cur[base] += 1
    if base in ("FFMA2", "FADD2", "FMUL2"):
        cur["FFMA2"] += 1
    if base in ("FFMA", "FADD", "FMUL"):
        cur["F32"] += 1
    if base in ("LDG", "STG"):
        cur[base] += 1
        if ".128" in op:
            cur[base + ".128"] += 1
    if base in ("HMMA", "IMMA", "DMMA", "QMMA", "UTCHMMA", "UTCQMMA", "UTCMMA", "HGMMA", "QGMMA"):
        cur["tensor"] += 1


def short(name):
    d = demangled.get(name, name)
    d = re.sub(r"plbm::\(anonymous namespace\)::|plbm::|\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    # cut the argument list: the first "(" outside the template brackets
    depth = 0
    for i, ch in enumerate(d):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return d[:i]
    return d


print(f"# {os.path.relpath(so, ROOT)}: cubins for {sorted(arch)}; {len(counts)} kernels; instruction counts per kernel (static SASS)")
print("# " + " ".join(f"{c:>7}" for c in COLS) + "  kernel")
tot = collections.Counter()
for (name, unit), c in sorted(counts.items(), key=lambda kv: (short(kv[0][0]), kv[0][1])):
    tot.update(c)
    print("  " + " ".join(f"{c.get(k, 0):>7}" for k in COLS) + "  " + short(name) + "  [" + unit + "]")
print("# " + " ".join(f"{tot.get(k, 0):>7}" for k in COLS) + "  TOTAL")
