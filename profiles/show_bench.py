"""Pretty-print the last JSON line of a bench log: python profiles/show_bench.py gpurun_out/bench.log"""
import json
import sys

line = [x for x in open(sys.argv[1]) if x.startswith("{")][-1]
d = json.loads(line)
print("value %.0f %s  ms/step %.3f  e2e %.0f  launches %s  clocks %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), d.get("clocks")))
for k in d.get("roofline_kernels", []):
    print("  %-18s %8.3f ms  share %.3f  frac %.4f" % (k["kernel"], k["ms_per_launch"], k["share"], k["frac"]))
nb = d.get("notebook_mode")
if nb:
    print("notebook mode: %.0f obj/s  %.3f ms  path frac %.4f" % (nb["objects_per_s"], nb["ms_per_step"], nb["roofline_path_frac"]))
    for k in nb["kernels"]:
        print("  %-18s %8.3f ms  frac %.4f" % (k["kernel"], k["ms_per_launch"], k["frac"]))
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"])
