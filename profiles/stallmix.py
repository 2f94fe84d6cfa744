#!/usr/bin/env python
"""Aggregate warp-stall sampling reasons of a kernel (ncu --page source --csv), overall and for FFMA instructions."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc = hdr.index("Source")
cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for _, h in cols}; ff = {h: 0 for _, h in cols}
for r in rows[2:]:
    try:
        vals = [int(r[i]) for i, _ in cols]
    except (ValueError, IndexError):
        break
    for (i, h), v in zip(cols, vals):
        tot[h] += v
        if "FFMA" in r[isrc]: ff[h] += v
s = sum(tot.values())
print("reason            all-instr   FFMA-only")
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print(f"  {h:18s} {100*v/s:6.2f}%   {100*ff[h]/s:6.2f}%")
