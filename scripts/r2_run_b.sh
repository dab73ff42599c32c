#!/bin/bash
# Round-2 GPU call B: ncu capture of the ORDERED kernel, thread-count timing, the tests that failed in call A, bench.
set -u
mkdir -p gpurun_out
for th in 128 256 512 1024; do timeout 120 python scripts/prof_ordered.py 200000 $th; done > gpurun_out/r2_ordered_threads.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ordered -s 1 -c 1 -f -o gpurun_out/r2_ordered python scripts/prof_ordered.py 200000 > gpurun_out/r2_ordered_ncu.log 2>&1
echo "ncu rc=$?"
timeout 900 python -m pytest tests/test_ordered_gpu.py -q -s > gpurun_out/r2_ordered.log 2>&1
echo "ordered rc=$?" | tee -a gpurun_out/r2_ordered.log
timeout 600 python -m pytest tests/test_hogwild_gpu.py tests/test_wavefront_gpu.py tests/test_parity_gpu.py -q -s > gpurun_out/r2_retest.log 2>&1
echo "retest rc=$?" | tee -a gpurun_out/r2_retest.log
timeout 600 python -X faulthandler bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
tail -n 5 gpurun_out/r2_ordered.log gpurun_out/r2_retest.log gpurun_out/r2_bench.err
cat gpurun_out/r2_ordered_threads.txt
grep "\[ordered" gpurun_out/r2_ordered.log
