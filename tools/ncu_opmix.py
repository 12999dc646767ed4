#!/usr/bin/env python3
"""Executed-instruction mix by SASS opcode from an .ncu-rep (source page, sass view).  python tools/ncu_opmix.py rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; mix = collections.Counter(); tot = 0
for r in rows:
    if r and r[0] == "Address": hdr = r; iE = hdr.index("Instructions Executed"); continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        ins = r[1].strip()
        if ins.startswith("@"): ins = ins.split(None, 1)[1]
        op = ins.split()[0].split(".")[0] if ins else "?"
        full = ins.split()[0]
        e = int(r[iE]); mix[full if len(sys.argv) > 2 else op] += e; tot += e
print("total", tot / 1e6, "M warp instructions")
for op, e in mix.most_common(45):
    print(f"{op:24s} {100 * e / tot:5.1f}%  {e / 1e6:8.2f} M")
