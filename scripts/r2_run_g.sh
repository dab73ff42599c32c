#!/bin/bash
# Round-2 GPU call G: ORDERED v5 (register-resident fast path) timing + ncu, the full GPU suite, bench.
set -u
mkdir -p gpurun_out
{
for v in 0 1 101 102 104 108 115; do timeout 120 python scripts/prof_ordered.py 200000 0 $v; done
for th in 256 512; do for v in 0 115; do timeout 120 python scripts/prof_ordered.py 200000 $th $v; done; done
timeout 120 python scripts/prof_ordered.py 1000209 0 0
timeout 120 python scripts/prof_ordered.py 1000209 0 1
} > gpurun_out/r2_ordered_v5.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ordered -s 1 -c 1 -f -o gpurun_out/r2_ordered_v5 python scripts/prof_ordered.py 200000 > gpurun_out/r2_ordered_ncu.log 2>&1
echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_gpu_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 600 python -X faulthandler bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cat gpurun_out/r2_ordered_v5.txt | cut -c1-120
tail -n 6 gpurun_out/r2_gpu_tests.log gpurun_out/r2_bench.err
grep "^\[" gpurun_out/r2_gpu_tests.log | cut -c1-200
