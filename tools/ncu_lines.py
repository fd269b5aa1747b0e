#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --print-source cuda,sass --csv` export:
share of warp-stall samples and of executed warp instructions, average active threads.  usage: ncu_lines.py file.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
        continue
    if hdr is None or len(r) <= iI or not r[0].isdigit():
        continue
    try:
        lines.append((int(r[0]), r[1], int(r[iS] or 0), int(r[iI] or 0), r[iT]))
    except ValueError:  # source lines with embedded quotes (inline asm in headers) break the CSV columns
        continue
ts, ti = sum(l[2] for l in lines), sum(l[3] for l in lines)
print(f"total samples {ts}  warp instructions {ti}")
for ln, src, s, i, t in lines:
    if 100 * s / max(ts, 1) >= thr or 100 * i / max(ti, 1) >= thr:
        print(f"{ln:>4} samples {100 * s / ts:5.1f}%  instr {100 * i / ti:5.1f}%  thr {t:>4}  {src.strip()[:120]}")
