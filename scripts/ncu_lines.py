"""Per-source-line summary of an ncu report: python scripts/ncu_lines.py REPORT.ncu-rep [top] > summary.txt
For every profiled kernel: the CUDA source lines with the most warp-stall samples and executed instructions
(ncu --page source --print-source cuda,sass; needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fn, fpath, hdr = None, None, None
agg = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))  # kernel -> (file, line) -> [samples, inst, text]
inst_total = defaultdict(int)
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
    elif row[0] == "Function Name":
        fn = row[1]
    elif row[0] == "Line No":
        hdr = row
        i_samp, i_inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and len(row) == len(hdr) and row[2] == "-":  # the CUDA line itself (its SASS rows follow)
        try:
            s, n = int(row[i_samp]), int(row[i_inst])
        except ValueError:
            continue
        a = agg[fn][(fpath, int(row[0]))]
        a[0] += s
        a[1] += n
        a[2] = row[1].strip()[:110]
        inst_total[fn] += n
for k, lines in agg.items():
    tot = sum(v[0] for v in lines.values()) or 1
    print("== %s   samples %d, warp instructions %d" % (k[:120], tot, inst_total[k]))
    for (f, ln), (s, n, txt) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %5.1f%% samples %5.1f%% inst  %s:%d  %s" % (100.0 * s / tot, 100.0 * n / max(inst_total[k], 1), f, ln, txt))
