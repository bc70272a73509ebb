"""Pretty-print a bench.py JSON line: python profiles/show_bench.py gpurun_out/bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d.pop("kernels", {})
print("value %.1f %s  ms/step %.2f  e2e %.1f  launches/step %s  tensor_frac %s  clocks %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches_per_step"),
    d.get("model_tensor_frac"), d.get("clocks")))
print("roofline", d.get("roofline"))
if "cpu_baseline" in d:
    print("cpu_baseline", d["cpu_baseline"])
for n, v in sorted(k.items(), key=lambda x: -x[1]["ms"]):
    print("%-20s %8.3f ms  share %.3f  %-6s %9.1f %-8s frac %.3f" % (n, v["ms"], v["share"], v["bound"], v["achieved"],
                                                                   v["unit"], v["frac"]))
