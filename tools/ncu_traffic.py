#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --page raw --csv` exports: DRAM bytes per launch of the dominant kernel, keyed
`<workload>/<precision>/<kernel>` (bench.py reads it for roofline.traffic).

    python tools/ncu_traffic.py <raw.csv> <workload> <precision> [<raw.csv> <workload> <precision> ...]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
table = json.load(open(path)) if os.path.exists(path) else {}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
args = sys.argv[1:]
for raw, workload, precision in zip(args[0::3], args[1::3], args[2::3]):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    per = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        m = re.search(r"(wn_layer_kernel|wn_gemm_kernel<\(?[a-zA-Z:() ]*(\d)\)?)", name)
        if not m:
            continue
        key = "wn_layer_kernel" if "wn_layer" in name else {"1": "wn_gemm_kernel<EPI_GATE>", "2": "wn_gemm_kernel<EPI_RESSKIP>"}.get(m.group(2))
        if key is None:
            continue
        b = sum(float(r[col[k]].replace(",", "")) * UNIT[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        t *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(units[col["gpu__time_duration.sum"]], 1.0)
        tensor = float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None
        per.setdefault(key, []).append((b, t, tensor))
    for key, v in per.items():
        table[f"{workload}/{precision}/{key}"] = {
            "dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "ncu_ms_per_launch": sum(x[1] for x in v) / len(v),
            "tensor_pipe_active_pct": None if v[0][2] is None else sum(x[2] for x in v) / len(v), "launches": len(v),
            "source": os.path.relpath(raw, ROOT)}
json.dump(table, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(table, indent=1, sort_keys=True))
