#!/usr/bin/env python
"""Dynamic instruction mix of every loop of one profiled kernel: `ncu -i X.ncu-rep --page source --csv > src.csv`, then
   python tools/ncu_loops.py src.csv <units per launch> [kernel#]
prints, per backward branch (loop), the warp instructions executed inside it per unit, by opcode, and its stall samples."""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[which]:starts[which + 1]]
print(blk[0][1])
hdr = blk[1]
iS, iE, iA, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("Warp Stall Sampling (All Samples)")
ins = [(int(r[iA], 16), r[iS].strip(), int(r[iE] or 0), int(r[iW] or 0)) for r in blk[2:] if len(r) > iE]
tot_s = sum(x[3] for x in ins)
tot_e = sum(x[2] for x in ins)
print("total executed per unit %.1f, samples %d" % (tot_e / units, tot_s))
loops = []
for a, t, e, w in ins:
    m = re.search(r"BRA\S*\s+(?:.*,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
base = ins[0][0]
for tgt, a in loops:
    body = [x for x in ins if tgt <= x[0] <= a]
    c = Counter()
    for x in body:
        op = x[1].split()[1] if x[1].startswith("@") else x[1].split()[0]
        c[".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDG", "STG", "LDS", "STS")) else op.split(".")[0]] += x[2]
    print("loop %x..%x: static %d, executed/unit %.1f (%.1f%%), samples %.1f%%" % (tgt - base, a - base, len(body), sum(x[2] for x in body) / units,
          100.0 * sum(x[2] for x in body) / tot_e, 100.0 * sum(x[3] for x in body) / max(tot_s, 1)))
    print("   ", {k: round(v / units, 1) for k, v in c.most_common(30)})
