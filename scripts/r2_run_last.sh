#!/bin/bash
# Round-2 last single-GPU check of the committed tree: the GPU suite, smoke, a default bench line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_last_tests.log 2>&1
echo "tests rc=$?"; tail -n 5 gpurun_out/r2_last_tests.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 1
timeout 600 python bench.py > gpurun_out/r2_last_bench.json 2> gpurun_out/r2_last_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_last_bench.json").read().strip().splitlines()[-1])
t = d["tolerance_mode"]
print("value %.4g ms %.4f frac %.3f e2e %.4g | ordered %.4g ms %.2f gap %.2g | cpu %.4g | launches %d" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], t["value"], t["ms_per_step"],
    t["parity"]["max_abs_gap"], d["cpu_baseline"]["value"], d["gpu_launches"]))
print({k: (v.get("value"), v.get("ms_per_step")) if isinstance(v, dict) else v for k, v in d["extra"].items()})
PY
