"""Per-source-line shared-memory wavefronts from an ncu report:
    python profiles/ncu_smem_lines.py report.ncu-rep kernel_regex [units] [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
norm = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) == len(hdr) and r[2] == "-":
        try:
            wf, ideal, inst = int(r[hdr.index("L1 Wavefronts Shared")]), int(r[hdr.index("L1 Wavefronts Shared Ideal")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        if wf:
            agg[(cur, int(r[0]))] = (wf, ideal, inst, r[1].strip())
tot = sum(v[0] for v in agg.values())
print("shared wavefronts %d (%.1f per unit)" % (tot, tot / norm))
for (f, ln), (wf, ideal, inst, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%8.1f %5.1f%% | ideal %7.1f | %s:%d  %s" % (wf / norm, 100.0 * wf / tot, ideal / norm, f, ln, src[:100]))
