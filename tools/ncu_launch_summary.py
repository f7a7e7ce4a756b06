#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / total / share."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
    m = re.search(r"wn_gemm_kernel<\(?[a-zA-Z:() ]*(\d)\)?, \(?[a-z]*\)?(\d)>", r[ki])
    if m:
        name = f"wn_gemm_kernel<EPI={m.group(1)},CG={m.group(2)}>"
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in agg.items():
    print(f"{k:44s} n={n:3d} total={t:9.3f} ms  avg={t / n:8.3f} ms  share={100 * t / tot:5.1f}%")
print(f"{'TOTAL':44s}       total={tot:9.3f} ms")
