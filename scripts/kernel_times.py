import json, sys
d = json.load(open(sys.argv[1]))
print("%.4g evals/s  %.4f ms/step  e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print({k: round(v["ms_per_iteration"], 4) for k, v in d["roofline"]["kernels"].items() if v["ms_per_iteration"] > 0})
