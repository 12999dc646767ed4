#!/usr/bin/env python3
"""Per-source-line instruction counts and stall samples from an .ncu-rep captured with --import-source on.
   python tools/ncu_lines.py rep.ncu-rep [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        lines.append(r)
iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iB = hdr.index("stall_barrier")
tot = sum(int(r[iE]) for r in lines); tots = sum(int(r[iS]) for r in lines)
print(f"total warp instructions {tot}, samples {tots}")
print("--- by instructions executed")
for r in sorted(lines, key=lambda r: -int(r[iE]))[:top]:
    print(f"{int(r[iE])*100/tot:5.1f}%  samp {int(r[iS])*100/max(tots,1):5.1f}%  L{r[0]:>5}: {r[1].strip()[:110]}")
print("--- by stall samples")
for r in sorted(lines, key=lambda r: -int(r[iS]))[:top//2]:
    print(f"{int(r[iS])*100/max(tots,1):5.1f}%  inst {int(r[iE])*100/tot:5.1f}%  L{r[0]:>5}: {r[1].strip()[:110]}")
