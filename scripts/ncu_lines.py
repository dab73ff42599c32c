"""Per-source-line stall samples of one ncu capture (development aid).
    python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None
rows = []
hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or not r[0].isdigit():
        continue
    try:
        samp = int(r[hdr["# Samples"]])
        inst = int(r[hdr["Instructions Executed"]])
    except (ValueError, KeyError):
        continue
    rows.append((samp, inst, cur, int(r[0]), r[1].strip()[:110]))
tot = sum(x[0] for x in rows) or 1
toti = sum(x[1] for x in rows) or 1
print("total samples %d, warp instructions %d" % (tot, toti))
for samp, inst, f, ln, src in sorted(rows, reverse=True)[:top]:
    print("%5.1f%% %6.1f%%i  %s:%d  %s" % (100.0 * samp / tot, 100.0 * inst / toti, f, ln, src))
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    print("---- lines %d..%d in file order ----" % (lo, hi))
    for samp, inst, f, ln, src in sorted(rows, key=lambda x: (x[2], x[3])):
        if lo <= ln <= hi and f.startswith("fm_ordered"):
            print("%5d samp %8d inst  %s:%d  %s" % (samp, inst, f, ln, src))
