#!/usr/bin/env python3
"""SASS executed under a source-line range: python tools/ncu_sass.py rep file lo hi [--list]"""
import csv, subprocess, sys, os, collections
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; line = None; mix = collections.Counter(); lst = []
for r in rows:
    if r and r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r and r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); continue
    if not hdr or len(r) != len(hdr): continue
    if r[0].isdigit(): line = int(r[0]); continue
    if r[0] == "" and cur == fname and line is not None and lo <= line <= hi:
        ins = r[3].strip()
        if not r[iE].isdigit(): continue
        e = int(r[iE])
        op = ins.split(None, 1)[1].split()[0] if ins.startswith("@") else ins.split()[0]
        mix[op] += e; lst.append((line, e, ins))
tot = sum(mix.values())
print("total", tot / 1e6, "M")
for op, e in mix.most_common(30): print(f"{op:22s} {e/1e6:8.2f} M {100*e/tot:5.1f}%")
if "--list" in sys.argv:
    for l, e, ins in lst: print(l, e, ins)
