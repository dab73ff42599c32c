"""Turn gpurun_out/*.ncu-rep / launch CSVs into the small tracked summaries under profiles/.

    python scripts/summarize_ncu.py <round-tag> <launch_csv | -> <ncu-rep> [traffic_json_name]
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
traffic_name = sys.argv[4] if len(sys.argv) > 4 else None
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = ["# %s -- ncu summary" % tag, ""]

# ---- launch list (cold-cache, serialised: compare SHARES, not absolutes) ----
rows = [r for r in csv.reader(open(launches_csv)) if r and not r[0].startswith("==")] if launches_csv != "-" else [[]]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rows[1:]:
    if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ci["Kernel Name"]].split("(")[0][:90]
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("u") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v_us
tot = sum(a[1] for a in agg.values()) or 1.0
if launches_csv != "-":
  lines += ["## launch list of `%s` (gpu__time_duration.sum, --clock-control none)" % os.path.basename(launches_csv), "",
          "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
  for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (name, n, us, 100 * us / tot, us / n))
  lines.append("")

# ---- full capture of the dominant kernel ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
m = {h[i]: (v[i], u[i]) for i in range(len(h))}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max"]
lines += ["## `ncu --set full` capture: %s" % os.path.basename(rep), "", "| metric | value | unit |", "|---|---:|---|"]
for k in want:
    if k in m:
        lines.append("| %s | %s | %s |" % (k, m[k][0][:110], m[k][1]))
stalls = sorted(((float(val.replace(",", "")), k) for k, (val, _) in m.items()
                 if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and val), reverse=True)[:6]
lines += ["", "top warp stall reasons (warps stalled per issue-active cycle):", ""]
for val, k in stalls:
    lines.append("* %.2f  %s" % (val, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))

def num(k):
    val, unit = m[k]
    x = float(val.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
lines += ["", "DRAM traffic of this launch: %.0f bytes (read %.0f + write %.0f)" % (
    dram, num("dram__bytes_read.sum"), num("dram__bytes_write.sum"))]
if traffic_name:
    json.dump({"dram_bytes_per_launch": dram, "source": os.path.basename(rep), "round": tag},
              open(os.path.join(out_dir, traffic_name), "w"))

# ---- hottest SASS lines ----
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
sr = list(csv.reader(src.splitlines()))
sh = sr[1]
si = {x: i for i, x in enumerate(sh)}
data = [r for r in sr[2:] if len(r) == len(sh)]
tot_s = sum(float(r[si["# Samples"]] or 0) for r in data) or 1
lines += ["", "hottest SASS instructions (share of warp-stall samples):", ""]
for r in sorted(data, key=lambda r: -float(r[si["# Samples"]] or 0))[:8]:
    lines.append("* %.1f%%  `%s`" % (100 * float(r[si["# Samples"]]) / tot_s, r[si["Source"]].strip()[:70]))
open(os.path.join(out_dir, tag + "_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
