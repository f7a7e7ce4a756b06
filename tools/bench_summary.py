"""One-line digest of bench.py JSON lines read from stdin."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    r = d["roofline"]
    ws = r.get("wavenet_stage", {})
    print(d["config"]["precision"], "cg", d["config"].get("tc_cta_group"),
          "value %.0f e2e %.0f ms/step %.2f | wn %.2f ms (gate %.2f, res/skip %.2f) | gate: %.0f alg TF, frac %.3f, exec %.3f |"
          % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["stages"]["wavenet"]["ms"], ws.get("gate_ms", 0), ws.get("resskip_ms", 0),
             r["achieved"], r["frac"], r["frac_executed"]),
          {k: round(v["ms"], 2) for k, v in d["stages"].items() if k != "wavenet"}, d["clocks"])
