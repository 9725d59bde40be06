"""Selected raw metrics per kernel from an ncu report:  python profiles/ncu_keys.py report.ncu-rep [tiles]"""
import csv, subprocess, sys
rep = sys.argv[1]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "smsp__inst_executed_op_shared_atom.sum",
        "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_shared_ld.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
for v in rows[2:]:
    d = dict(zip(h, v))
    print("----")
    for k in keys:
        if k in d:
            x = d[k]
            extra = ""
            if tiles and k.endswith(".sum") and "pct" not in k and "time" not in k and "dram" not in k:
                try:
                    extra = "   (%.1f per tile)" % (float(x.replace(",", "")) / tiles)
                except ValueError:
                    pass
            print("%-90s %s%s" % (k, x, extra))
