import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d["config"]["precision"], "cg", d["config"].get("tc_cta_group"), "value %.0f e2e %.0f ms/step %.2f | wn %.2f ms, alg TF %.0f, frac_exec %.3f |" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["stages"]["wavenet"]["ms"], d["roofline"]["achieved"], d["roofline"]["frac_executed"]), {k:round(v["ms"],2) for k,v in d["stages"].items() if k!="wavenet"}, d["clocks"])
