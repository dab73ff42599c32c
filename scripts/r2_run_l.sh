#!/bin/bash
# Round-2 GPU call L: ORDERED v10 (v9 helpers + 8-segment shuffle chain with identity-padded pairs): phase timers, ncu, tests.
set -u
mkdir -p gpurun_out
{
for v in 0 132 2 147; do timeout 60 python scripts/prof_ordered.py 200000 0 $v; done
timeout 60 python scripts/prof_ordered.py 1000209 0 0
} > gpurun_out/r2_ordered_v10.txt 2>&1
cat gpurun_out/r2_ordered_v10.txt | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ordered -s 1 -c 1 -f -o gpurun_out/r2_ordered_v10 python scripts/prof_ordered.py 200000 > gpurun_out/r2_ordered_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_ordered_gpu.py -m gpu -q -s -x > gpurun_out/r2_gpu_tests_ordered.log 2>&1
echo "ordered tests rc=$?"
tail -n 5 gpurun_out/r2_gpu_tests_ordered.log | cut -c1-300
