#!/bin/bash
# Round-2 GPU call I: ORDERED v7 (warp-specialised driver, segmented chain, one-hot path): timing, ncu, tests, bench.
set -u
mkdir -p gpurun_out
{
for v in 0 2 1 101 102 104 108 115; do timeout 60 python scripts/prof_ordered.py 200000 0 $v; done
timeout 60 python scripts/prof_ordered.py 1000209 0 0
timeout 60 python scripts/prof_ordered.py 1000209 0 2
} > gpurun_out/r2_ordered_v7.txt 2>&1
cat gpurun_out/r2_ordered_v7.txt | cut -c1-100
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ordered -s 1 -c 1 -f -o gpurun_out/r2_ordered_v7 python scripts/prof_ordered.py 200000 > gpurun_out/r2_ordered_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_ordered_gpu.py tests/test_wavefront_gpu.py -m gpu -q -s -x > gpurun_out/r2_gpu_tests_ordered.log 2>&1
echo "ordered tests rc=$?"
tail -n 5 gpurun_out/r2_gpu_tests_ordered.log
timeout 200 python -m pytest tests/test_hogwild_gpu.py -m gpu -q -s -k peer > gpurun_out/r2_gpu_tests_peer.log 2>&1
echo "peer tests rc=$?"
tail -n 3 gpurun_out/r2_gpu_tests_peer.log
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_ordered_gpu.py --deselect tests/test_wavefront_gpu.py -k "not peer" > gpurun_out/r2_gpu_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 600 python -X faulthandler bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
tail -n 6 gpurun_out/r2_gpu_tests.log gpurun_out/r2_bench.err
grep "^\[" gpurun_out/r2_gpu_tests.log gpurun_out/r2_gpu_tests_ordered.log | cut -c1-200
