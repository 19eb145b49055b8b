#!/usr/bin/env python
"""Per-source-line instruction counts of one profiled kernel.

Joins `ncu --page source --csv` (per-SASS-instruction counters) with `nvdisasm -g` line info of the
in-tree library, and prints warp instructions executed + stall samples per file:line, so a profile can
be read as "which source statement costs how many issue slots per record".

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_canon_w2 [--per N] [--top K]
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "circkit_b200", "libcirckit_b200.so")


def line_map(kernel_pat: str, lib: str = LIB):
    """{function name: {offset: (file, line)}} for functions matching kernel_pat."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], check=True, capture_output=True,
                             text=True).stdout
        cur, loc = None, ("?", 0)
        for ln in txt.splitlines():
            m = re.match(r"\s*\.type\s+(\S+),@function", ln)
            if m:
                cur = m.group(1) if re.search(kernel_pat, m.group(1)) else None
                if cur:
                    # out-of-line device functions ($kernel$callee) share the kernel's section and address space
                    sub = re.match(r"\$([^$]+)\$", cur)
                    if sub:
                        cur = sub.group(1)
                    out.setdefault(cur, {})
                continue
            if cur is None:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out[cur][int(m.group(1), 16)] = (loc, m.group(2).strip())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--per", type=float, default=1.0, help="divide instruction counts by this (e.g. records)")
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--sass", action="store_true", help="also print the hottest SASS instructions")
    ap.add_argument("--launch", type=int, default=0, help="which profiled launch of the kernel (0 = first)")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], check=True, capture_output=True,
                         text=True).stdout
    # the CSV holds one block per profiled launch: "Kernel Name",<name> / header / rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    sel = [b for b in blocks if re.search(a.kernel, b["name"])]
    if not sel:
        sys.exit("no launch of %s in %s (%s)" % (a.kernel, a.rep, [b["name"] for b in blocks]))
    b = sel[a.launch]
    hdr = b["hdr"]
    iA, iS, iN, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    maps = line_map(a.kernel)
    # pick the function whose instruction count matches
    fn = None
    for name, m in maps.items():
        if len(m) == len(b["rows"]):
            fn = name
    if fn is None:
        sys.exit("no function with %d instructions among %s" % (len(b["rows"]), {k: len(v) for k, v in maps.items()}))
    m = maps[fn]
    base = int(b["rows"][0][iA], 16)
    per_line = collections.defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    sass = []
    for r in b["rows"]:
        off = int(r[iA], 16) - base
        loc, text = m.get(off, (("?", 0), "?"))
        ex, sm = int(r[iE]), int(r[iN])
        per_line[loc][0] += ex
        per_line[loc][1] += sm
        tot_i += ex
        tot_s += sm
        sass.append((ex, sm, off, loc, r[iS].strip()))
    print("# %s  (%s)" % (b["name"], fn))
    print("# warp instructions executed: %d  (%.1f per unit), samples %d" % (tot_i, tot_i / a.per, tot_s))
    print("%-22s %12s %7s %7s" % ("file:line", "inst/unit", "inst%", "samp%"))
    for loc, (ex, sm) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print("%-22s %12.2f %6.1f%% %6.1f%%" % ("%s:%d" % loc, ex / a.per, 100.0 * ex / max(tot_i, 1), 100.0 * sm / max(tot_s, 1)))
    if a.sass:
        print("# hottest SASS")
        for ex, sm, off, loc, text in sorted(sass, key=lambda t: -t[1])[: a.top]:
            print("%6x %-20s %10.2f %6.1f%%  %s" % (off, "%s:%d" % loc, ex / a.per, 100.0 * sm / max(tot_s, 1), text))


if __name__ == "__main__":
    main()
