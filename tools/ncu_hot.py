#!/usr/bin/env python
"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv [--kernel-name ...] > src.csv`:
stall samples and executed instructions by opcode and by SASS window.   python tools/ncu_hot.py src.csv [window] [kernel#]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 60
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[which]:starts[which + 1]]
print(blk[0][1])
hdr = blk[1]
iS, iA, iE = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(r[iS].strip(), int(r[iA] or 0), int(r[iE] or 0)) for r in blk[2:] if len(r) > iE]
tot, tote = sum(d[1] for d in data), sum(d[2] for d in data)
print("samples", tot, "warp instructions", tote, "SASS lines", len(data))
cs, ce = Counter(), Counter()
for s, a, e in data:
    t = s.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    cs[op] += a; ce[op] += e
for op, a in cs.most_common(16):
    print("%-10s samples %6d (%4.1f%%)  executed %10d (%4.1f%%)" % (op, a, 100 * a / tot, ce[op], 100 * ce[op] / tote))
seg = sorted(((sum(x[1] for x in data[i:i + W]), sum(x[2] for x in data[i:i + W]), i) for i in range(0, len(data), W)), reverse=True)
for a, e, i in seg[:12]:
    print("sass %5d..%5d samples %6d (%4.1f%%) executed %10d (%4.1f%%)" % (i, i + W, a, 100 * a / tot, e, 100 * e / tote))
