#!/usr/bin/env python
"""Executed SASS of one profiled kernel in address order, with per-instruction executed counts and stall samples.
   python tools/ncu_sass.py rep kernel_regex per [min_count]"""
import csv, io, subprocess, sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
import ncu_lines as NL
rep, pat, per = sys.argv[1], sys.argv[2], float(sys.argv[3])
minc = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
import re
b = [b for b in blocks if re.search(pat, b["name"])][int(sys.argv[5]) if len(sys.argv) > 5 else 0]
hdr = b["hdr"]; iA, iS, iE, iN = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
import os
maps = NL.line_map(os.environ.get("MANGLED", pat))
cands = [v for v in maps.values() if len(v) == len(b["rows"])] or [v for v in maps.values() if len(v) <= len(b["rows"])]
m = max(cands, key=len)
base = int(b["rows"][0][iA], 16)
tot = 0
for r in b["rows"]:
    off = int(r[iA], 16) - base
    loc, _ = m.get(off, (("?", 0), ""))
    c = int(r[iE]) / per
    tot += c
    if c >= minc:
        print("%5x %-22s %8.2f %5d  %s" % (off, "%s:%d" % loc, c, int(r[iN]), r[iS].strip()))
print("total per unit: %.1f" % tot)
