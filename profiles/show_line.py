"""One-line digest of a bench.py JSON line:  python bench.py ... | python profiles/show_line.py label"""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    ks = {k["kernel"].split(" ")[0]: round(k["ms_per_step"], 3) for k in d.get("roofline_kernels", [])}
    nb = d.get("notebook_mode", {})
    print(sys.argv[1] if len(sys.argv) > 1 else "", "N=%d obj/s %.0f ms %.3f frac %.4f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"]), ks,
          "| notebook %.0f ms %.3f" % (nb.get("objects_per_s", 0), nb.get("ms_per_step", 0)),
          {k["kernel"].split(" ")[0]: round(k["ms_per_launch"], 3) for k in nb.get("kernels", [])},
          "| e2e %.0f bits %.0f" % (d.get("e2e", {}).get("value", 0), d.get("e2e_bitmask", {}).get("value", 0)),
          "| cfg4", {k: round(v) for k, v in d.get("cfg4_ablation_sweep", {}).items() if k.endswith("per_s")},
          "| cfg5 %.0f" % d.get("cfg5_variable_sparse", {}).get("objects_per_s", 0),
          "| 16bit %.0f" % d.get("full_16bit_range", {}).get("objects_per_s", 0),
          "| basic frac", d.get("roofline_basic_block", {}).get("frac"), d.get("gather_check"), d["config"].get("collective", "")[:60])
