"""profiles/make_traffic.py REPORT.ncu-rep OUT.json -- DRAM bytes, duration and shared-memory pipe load per kernel of
one step from a `--set full` capture (what bench.py reads as roofline.traffic)."""
import csv
import json
import subprocess
import sys

rep, dst = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U, D = rows[0], rows[1], rows[2:]


def val(r, n):
    v = float(r[H.index(n)].replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "ns": 1e-6}.get(U[H.index(n)], 1)


K, tot = {}, 0.0
for r in D:
    name = r[H.index("Kernel Name")].replace("void imfeat::", "").replace("imfeat::", "").split("(")[0]
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    wf = float(r[H.index("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")].replace(",", ""))
    cyc = float(r[H.index("sm__cycles_elapsed.avg")].replace(",", ""))
    K[name] = {"dram_bytes": rd + wr, "read": rd, "write": wr, "ms": val(r, "gpu__time_duration.sum"),
               "shared_wavefronts_per_sm_clock": wf / 148 / cyc}
    tot += rd + wr
alg = 10000 * (2 * 4096 * 12 + 4096 * 12 + 8 * 60 * 12)
json.dump({"source": "ncu --set full capture of one step inside `python bench.py --steps 2 --warmup 3 --no-cpu --no-side` "
                     "(10,000 objects 64x64x12 + masks, all blocks); summary in profiles/r2_ncu_bench_full_10k_summary.txt",
           "kernels": K, "step_total_bytes": tot, "algorithmic_bytes": alg, "ratio": tot / alg}, open(dst, "w"), indent=1)
print({k: (round(v["ms"], 3), round(v["dram_bytes"] / 1e9, 3)) for k, v in K.items()}, round(tot / 1e9, 3), round(tot / alg, 3))
