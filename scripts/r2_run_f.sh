#!/bin/bash
# Round-2 GPU call F: ORDERED v5 (register-resident fast path, dense write-back): phase timing, thread counts, tests.
set -u
mkdir -p gpurun_out
{
for v in 0 101 102 104 108 115; do timeout 120 python scripts/prof_ordered.py 200000 0 $v; done
for th in 256 512; do for v in 0 101 115; do timeout 120 python scripts/prof_ordered.py 200000 $th $v; done; done
timeout 120 python scripts/prof_ordered.py 200000 0 1
timeout 120 python scripts/prof_ordered.py 1000209 0 0
timeout 120 python scripts/prof_ordered.py 1000209 512 0
} > gpurun_out/r2_ordered_v5.txt 2>&1
timeout 900 python -m pytest tests/test_ordered_gpu.py -q -s > gpurun_out/r2_gpu_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
cat gpurun_out/r2_ordered_v5.txt | cut -c1-100
tail -n 6 gpurun_out/r2_gpu_tests.log
grep "^\[" gpurun_out/r2_gpu_tests.log | cut -c1-160
