"""Static evidence from the built library: per kernel, registers / shared memory (ptxas -v) and the counts of the
SASS mnemonics the design relies on (packed FFMA2 / FMUL2 / FADD2, MUFU, LDGSTS = cp.async, SYNCS = mbarrier,
REDG = red.global, LDGMC = multimem.ld_reduce, SHFL, CREDUX = redux.sync).
    python tools/sass_summary.py > profiles/<round>_sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "eogs2_b200" / "libeogs_raster.so"
KEYS = ["FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "LDGSTS", "SYNCS", "REDG", "RED", "ATOMG", "LDGMC", "SHFL", "CREDUX", "VOTE",
        "LDS", "STS", "LDG", "STG", "BAR"]

sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
kern = OrderedDict()
cur = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        kern[cur][m.group(1)] += 1
        kern[cur]["_total"] += 1
res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
usage = {}
name = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*SHARED:(\d+)", ln)
    if m and name:
        usage[name] = (int(m.group(1)), int(m.group(2)))
demangle = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# {LIB.name}: SASS mnemonic counts per kernel (static), registers and static shared memory")
for (mangled, c), nice in zip(kern.items(), demangle):
    if not nice.replace("void ", "").startswith("eogs::"):
        continue
    short = re.sub(r"\(.*", "", nice).replace("void ", "")
    reg, smem = usage.get(mangled, ("?", "?"))
    parts = [f"{k}={c[k]}" for k in KEYS if c[k]]
    print(f"{short:55s} regs={reg:>3} smem={smem:>6}  instrs={c['_total']:>5}  " + " ".join(parts))
