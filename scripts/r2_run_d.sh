#!/bin/bash
# Round-2 GPU call D: ORDERED v3 (tile-end write-back) timing experiments + ncu, tests, sweep with the bias ramp, bench.
set -u
mkdir -p gpurun_out
{
timeout 120 python scripts/prof_ordered.py 200000 0 0
timeout 120 python scripts/prof_ordered.py 200000 0 101
timeout 120 python scripts/prof_ordered.py 200000 64 0
timeout 120 python scripts/prof_ordered.py 200000 32 0
timeout 120 python scripts/prof_ordered.py 1000209 0 0
} > gpurun_out/r2_ordered_v3.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ordered -s 1 -c 1 -f -o gpurun_out/r2_ordered_v3 python scripts/prof_ordered.py 200000 > gpurun_out/r2_ordered_ncu.log 2>&1
echo "ncu rc=$?"
timeout 900 python -m pytest tests/test_ordered_gpu.py tests/test_hogwild_gpu.py tests/test_cli_gpu.py -q -s > gpurun_out/r2_gpu_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 400 python scripts/sweep_hogwild.py --out gpurun_out/r2_sweep.json > gpurun_out/r2_sweep.log 2>&1
echo "sweep rc=$?"
timeout 600 python -X faulthandler bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
cat gpurun_out/r2_ordered_v3.txt
tail -n 6 gpurun_out/r2_gpu_tests.log gpurun_out/r2_bench.err
grep "^\[" gpurun_out/r2_gpu_tests.log | cut -c1-200
tail -n 17 gpurun_out/r2_sweep.log
