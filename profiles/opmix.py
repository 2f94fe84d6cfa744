#!/usr/bin/env python
"""Per-opcode executed warp-instruction mix of a kernel from `ncu --page source --csv` output (stdin or file)."""
import csv, sys, io, re
from collections import defaultdict
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = defaultdict(lambda: [0, 0]); tot = 0; tots = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    if not m: continue
    op = m.group(2)
    base = op.split(".")[0]
    if base in ("LDS", "STS", "LDG", "STG"):
        base = ".".join(op.split(".")[:1]) + ("." + [p for p in op.split(".") if p in ("64", "128")][0] if any(p in ("64", "128") for p in op.split(".")) else "")
    try:
        n = int(r[iex]); s = int(r[ismp])
    except ValueError:
        break   # next section (another view of the same kernel)
    ops[base][0] += n; ops[base][1] += s; tot += n; tots += s
print(f"total warp instructions executed: {tot:,}   samples: {tots:,}")
for k, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"  {k:14s} {n:>14,} {100*n/tot:6.2f}%   stall samples {100*s/max(tots,1):6.2f}%")
