#!/bin/bash
# Round-2 GPU call E: ORDERED v4 phase timing (debug bits), helper-warp counts, barrier/store microbench, tests, bench.
set -u
mkdir -p gpurun_out
timeout 60 ./scripts/micro/bar_store_bench > gpurun_out/r2_bar_store_bench.txt 2>&1
{
for v in 0 101 102 104 108 103 107 115; do timeout 120 python scripts/prof_ordered.py 200000 0 $v; done
for th in 256 512; do timeout 120 python scripts/prof_ordered.py 200000 $th 0; done
timeout 120 python scripts/prof_ordered.py 1000209 0 0
timeout 120 python scripts/prof_ordered.py 1000209 512 0
} > gpurun_out/r2_ordered_v4.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ordered -s 1 -c 1 -f -o gpurun_out/r2_ordered_v4 python scripts/prof_ordered.py 200000 > gpurun_out/r2_ordered_ncu.log 2>&1
echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_gpu_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 600 python -X faulthandler bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
cat gpurun_out/r2_ordered_v4.txt
tail -n 6 gpurun_out/r2_gpu_tests.log gpurun_out/r2_bench.err
grep "^\[" gpurun_out/r2_gpu_tests.log | cut -c1-200
cat gpurun_out/r2_bar_store_bench.txt
