import json,sys
d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k in d["kernels"]: print("  %-18s %.4f ms  %s" % (k["kernel"], k["ms"], ("%.0f %s frac %.3f" % (k["achieved"], k["unit"], k["frac"])) if "frac" in k else ""))
