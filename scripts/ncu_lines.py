#!/usr/bin/env python3
"""Per CUDA source line: stall samples (with the two top stall reasons) and warp instructions of one .ncu-rep (compiled
with -lineinfo, captured with --import-source on).  Usage: python scripts/ncu_lines.py rep.ncu-rep [top_n]"""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
hdr, fname, acc = None, "", []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if not hdr or r[0] in ("", "...") or len(r) < len(hdr):
        continue
    idx = {h: i for i, h in enumerate(hdr) if h not in ("Source",)}
    try:
        smp, n = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]])
    except ValueError:
        continue
    stalls = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for h, i in idx.items() if h.startswith("stall_") and "Not Issued" not in h), reverse=True)
    acc.append((smp, n, fname, r[0], r[1][:110].strip(), ", ".join(f"{h[6:]} {v}" for v, h in stalls[:2] if v)))
tot = sum(a[0] for a in acc) or 1
totn = sum(a[1] for a in acc) or 1
print(f"# {rep}: {tot} samples, {totn} warp instructions")
for smp, n, f, line, src, st in sorted(acc, reverse=True)[:top]:
    print(f"{100.0 * smp / tot:6.2f}% smp {100.0 * n / totn:6.2f}% inst  {f}:{line:>4s}  {src}   [{st}]")
