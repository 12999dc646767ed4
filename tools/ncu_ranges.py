#!/usr/bin/env python3
"""Instruction share per source-line range (phase) from an .ncu-rep.  python tools/ncu_ranges.py rep a:b[:name] ..."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit(): lines.append(r)
iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
tot = sum(int(r[iE]) for r in lines); ts = sum(int(r[iS]) for r in lines)
for spec in sys.argv[2:]:
    parts = spec.split(":"); a, b = int(parts[0]), int(parts[1]); name = parts[2] if len(parts) > 2 else ""
    e = sum(int(r[iE]) for r in lines if a <= int(r[0]) <= b); s = sum(int(r[iS]) for r in lines if a <= int(r[0]) <= b)
    print(f"L{a:5d}-{b:5d} {name:28s} inst {100*e/tot:5.1f}%  ({e/1e6:7.1f}M)   samples {100*s/ts:5.1f}%")
print("total", tot/1e6, "M")
