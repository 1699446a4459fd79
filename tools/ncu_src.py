#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: stall mix, hottest SASS lines, opcode histogram.
usage: ncu -i rep --page source --csv --kernel-name regex:<k> > x.csv ; python tools/ncu_src.py x.csv [ntop]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
print("samples", tot, "warp-instr", sum(int(r[iI]) for r in data), "sass lines", len(data))
agg = {s: sum(int(r[hdr.index(s)] or 0) for r in data) for s in stalls}
print("stalls:", [(k, v, "%.0f%%" % (100.0 * v / max(tot, 1))) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
for r in sorted(data, key=lambda r: -int(r[iS]))[:ntop]:
    top = sorted(((s, int(r[hdr.index(s)] or 0)) for s in stalls), key=lambda x: -x[1])[0]
    print("%6s %10s  %-70s %s" % (r[iS], r[iI], r[isrc].strip()[:70], top))
c = Counter()
for r in data:
    t = r[isrc].split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    c[op.split(".")[0]] += int(r[iI])
print(c.most_common(30))
