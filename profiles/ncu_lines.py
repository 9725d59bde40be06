"""Per-source-line instruction and stall-sample totals from an ncu report.

    python profiles/ncu_lines.py report.ncu-rep kernel_regex [tiles_x_warps] [top]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
norm = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, lines, tot_i, tot_s = None, [], 0, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] and r[0].isdigit() and len(r) >= 8 and r[2] == "-":
        try:
            inst, samp = int(r[7]), int(r[6])
        except ValueError:
            continue
        lines.append((inst, samp, cur_file, int(r[0]), r[1].strip()))
        tot_i += inst
        tot_s += samp
print("total instructions %d  (%.1f per unit)  samples %d" % (tot_i, tot_i / norm, tot_s))
key = (lambda t: t[1]) if (len(sys.argv) > 5 and sys.argv[5] == "stall") else (lambda t: t[0])
for inst, samp, f, ln, src in sorted(lines, key=key, reverse=True)[:top]:
    print("%8.1f %5.1f%% inst | %5.1f%% stall | %s:%d  %s" % (inst / norm, 100.0 * inst / tot_i,
                                                           100.0 * samp / max(tot_s, 1), f, ln, src[:95]))
