"""SASS instructions of one ncu capture in address order with stall samples and execution counts
(development aid).   python scripts/ncu_sass.py x.ncu-rep [min_samples] > out.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] in ("Address", "Line No") or "Source" in r[:3] and hdr is None:
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None:
        continue
    try:
        samp = int(r[hdr["# Samples"]]); inst = int(r[hdr["Instructions Executed"]])
    except Exception:
        continue
    print("%6d %8d  %s" % (samp, inst, r[hdr["Source"]]))
