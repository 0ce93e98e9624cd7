#!/usr/bin/env python
"""Summarises `ncu --page source --csv` output: executed warp-instructions by opcode, stall samples by reason,
and the hottest SASS lines.  usage: ncu_src_summary.py file.csv [cells]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
ops, stall = Counter(), Counter()
total = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    try:
        n = int(float(r[col["Instructions Executed"]]))
    except ValueError:
        continue
    toks = src.split()
    op = toks[0] if toks and not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("IMAD", "LDS", "STG", "LDG", "ATOM", "RED")) and "." in op else "")
    ops[op] += n
    total += n
    samples = int(float(r[col["# Samples"]] or 0))
    lines.append((samples, n, src))
    for h, i in col.items():
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                stall[h] += int(float(r[i] or 0))
            except ValueError:
                pass
print(f"total warp-instructions executed: {total:,}" + (f"  = {total * 32 / cells:.1f} lane-instr per cell" if cells else ""))
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n:>14,} {100 * n / total:5.1f}%" + (f"  {n * 32 / cells:6.2f}/cell" if cells else ""))
ts = sum(stall.values())
print("stall samples:", ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in stall.most_common(10)))
print("hottest lines by samples:")
for s, n, src in sorted(lines, reverse=True)[:25]:
    print(f"  {s:6d} {n:>12,}  {src[:100]}")
