#!/usr/bin/env python
"""Top stalled SASS instructions per kernel from `ncu -i X.ncu-rep --page source --csv [--print-source sass]`."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern, hdr, rows = None, None, []


def flush():
    if not rows:
        return
    si = hdr.index("# Samples")
    tot = sum(int(r[si] or 0) for r in rows)
    print(f"== {kern[:90]}  total samples {tot}")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][si] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
        print(f"  {i:5d} {int(r[si]):7d} {100 * int(r[si]) / max(tot, 1):5.1f}%  {r[1].strip()[:80]:80s} {st}")


for r in csv.reader(open(path)):
    if r and r[0] == "Kernel Name":
        flush()
        kern, rows, hdr = r[1], [], None
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) == len(hdr):
        rows.append(r)
flush()
