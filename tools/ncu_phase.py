#!/usr/bin/env python3
"""Instruction / stall-sample share per (file, line range) from an .ncu-rep captured with --import-source on.
   python tools/ncu_phase.py rep.ncu-rep [file:lo-hi:name ...]   (no specs: per-file totals + top lines with file names)"""
import csv, subprocess, sys, os
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; recs = []
for r in rows:
    if r and r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit() and len(r) == len(hdr):
        recs.append((cur, int(r[0]), r[1].strip(), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])))
tot = sum(x[3] for x in recs); ts = sum(x[4] for x in recs)
print(f"total warp instructions {tot/1e6:.1f} M, samples {ts}")
specs = sys.argv[2:]
if not specs:
    files = {}
    for f, l, s, e, sm in recs:
        a = files.setdefault(f, [0, 0]); a[0] += e; a[1] += sm
    for f, (e, sm) in files.items(): print(f"{f:24s} inst {100*e/tot:5.1f}%  samples {100*sm/ts:5.1f}%")
    for f, l, s, e, sm in sorted(recs, key=lambda x: -x[3])[:60]:
        print(f"{100*e/tot:5.1f}% samp {100*sm/ts:5.1f}%  {f}:{l}: {s[:100]}")
for spec in specs:
    f, rng, *name = spec.split(":"); lo, hi = map(int, rng.split("-"))
    e = sum(x[3] for x in recs if x[0] == f and lo <= x[1] <= hi); sm = sum(x[4] for x in recs if x[0] == f and lo <= x[1] <= hi)
    print(f"{f}:{lo}-{hi} {(name[0] if name else ''):26s} inst {100*e/tot:5.1f}% ({e/1e6:7.1f} M)  samples {100*sm/ts:5.1f}%")
