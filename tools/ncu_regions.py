"""Group the SASS of one kernel (ncu --page source --csv export) into regions of equal execution count and print
instructions, stall-sample share and the top stall reasons per region.  usage: ncu_regions.py <source.csv> [min Minstr]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
# several kernels in one export: sections start with a "Kernel Name" row; take the first (or the one whose name contains argv[3])
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = 0
if len(sys.argv) > 3:
    pick = next((k for k, i in enumerate(starts) if sys.argv[3] in rows[i][1]), 0)
lo = starts[pick]
hi = starts[pick + 1] if pick + 1 < len(starts) else len(rows)
print(rows[lo][1][:100])
hdr, data = rows[lo + 1], [r for r in rows[lo + 2:hi] if len(r) > 5]
ix = {h: i for i, h in enumerate(hdr)}
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
tot_s = sum(int(r[ix['# Samples']] or 0) for r in data)
tot_i = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and '(Not' not in h]
seg = []
for n, r in enumerate(data):
    e = int(r[ix['Instructions Executed']] or 0) / 1e6
    op = [o for o in r[ix['Source']].split() if not o.startswith('@')]
    seg.append((n, e, int(r[ix['# Samples']] or 0), op[0] if op else '', r))
groups = []
for t in seg:
    if groups and abs(groups[-1][-1][1] - t[1]) <= 0.03 * max(t[1], 1e-9):
        groups[-1].append(t)
    else:
        groups.append([t])
print(f"total {tot_i/1e6:.0f}M warp instructions, {tot_s} samples")
for g in groups:
    if g[0][1] * len(g) < thr:
        continue
    c = Counter(x[3].split('.')[0] for x in g)
    tot = {s: 0 for s in stalls}
    for x in g:
        for s in stalls:
            tot[s] += int(x[4][ix[s]] or 0)
    T = sum(tot.values()) or 1
    top = ', '.join(f"{k[6:]}:{100*v/T:.0f}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:4])
    print(f"[{g[0][0]}-{g[-1][0]}] n={len(g)} exec={g[0][1]:.2f}M inst={len(g)*g[0][1]:.0f}M ({100*len(g)*g[0][1]*1e6/tot_i:.1f}%) "
          f"samples={100*sum(x[2] for x in g)/tot_s:.1f}% | {top} |", dict(c.most_common(6)))
