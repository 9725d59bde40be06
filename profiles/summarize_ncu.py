"""Summarise an ncu report (one --set full capture) into a small text table.

    python profiles/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x_summary.txt
"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "smsp__inst_executed_op_shared_atom.sum",
]
PAT = re.compile(r"^(sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|fp64|xu|adu|cbu|uniform)\.avg\.pct_of_peak_sustained_active"
                 r"|smsp__average_warps_issue_stalled_(barrier|long_scoreboard|short_scoreboard|math_pipe_throttle|"
                 r"mio_throttle|lg_throttle|wait|not_selected|branch_resolving|no_instruction)_per_issue_active\.ratio)$")
print("# source: %s (ncu --set full --clock-control none, values per launch)" % rep)
names = [r[hdr.index("Kernel Name")].replace("void imfeat::", "").split("(")[0][:26] for r in data]
print("%-86s %s" % ("metric", " | ".join(n.rjust(26) for n in names)))
for i, h in enumerate(hdr):
    if h in KEEP[1:] or PAT.match(h):
        print("%-86s %s  %s" % (h, " | ".join(r[i][:26].rjust(26) for r in data), units[i]))
