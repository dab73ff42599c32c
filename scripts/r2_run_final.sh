#!/bin/bash
# Round-2 final single-GPU call on the committed tree: GPU suite, smoke, bench (both arms), the ncu launch list of
# the bench command and one --set full capture each of the HOGWILD C2 epoch kernel and the ORDERED epoch kernel.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_final_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_final_tests.log
tail -n 4 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
echo "bench ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc=$?"; tail -n 3 gpurun_out/r2_final_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-extras --no-parity --no-cpu-baseline > gpurun_out/r2_final_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowlane -s 2 -c 1 -f -o gpurun_out/r2_final_c2_epoch \
  python scripts/prof_epochs.py c2 4 > gpurun_out/r2_final_ncu_c2.log 2>&1
echo "c2 capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ordered -s 1 -c 1 -f -o gpurun_out/r2_final_ordered \
  python scripts/prof_ordered.py 200000 > gpurun_out/r2_final_ncu_ordered.log 2>&1
echo "ordered capture rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
t = d["tolerance_mode"]
print("value %.4g ms %.4f frac %.3f e2e %.4g | ordered value %.4g ms %.2f e2e %.4g gap %.2g | cpu %.4g" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], t["value"], t["ms_per_step"],
    t["e2e"]["value"], t["parity"]["max_abs_gap"], d["cpu_baseline"]["value"]))
print({k: (v.get("value"), v.get("ms_per_step")) if isinstance(v, dict) else v for k, v in d["extra"].items()})
print(d["clocks"], "wall", d["timed_region_wall_s"])
PY
grep "^\[" gpurun_out/r2_final_tests.log | cut -c1-180 | tail -n 30
