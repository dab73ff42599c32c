#!/bin/bash
# Round-2 GPU call C: ORDERED v2 timing + ncu, TMA reduction microbench, full GPU suite, hogwild sweep, bench.
set -u
mkdir -p gpurun_out
timeout 120 python scripts/prof_ordered.py 200000 > gpurun_out/r2_ordered_v2.txt 2>&1
timeout 120 python scripts/prof_ordered.py 1000209 >> gpurun_out/r2_ordered_v2.txt 2>&1
timeout 60 ./scripts/micro/tma_red_bench > gpurun_out/r2_tma_red_bench.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ordered -s 1 -c 1 -f -o gpurun_out/r2_ordered_v2 python scripts/prof_ordered.py 200000 > gpurun_out/r2_ordered_ncu.log 2>&1
echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 400 python scripts/sweep_hogwild.py --out gpurun_out/r2_sweep.json > gpurun_out/r2_sweep.log 2>&1
echo "sweep rc=$?"
timeout 600 python -X faulthandler bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
cat gpurun_out/r2_ordered_v2.txt gpurun_out/r2_tma_red_bench.txt
tail -n 6 gpurun_out/r2_gpu_tests.log gpurun_out/r2_bench.err
grep "^\[" gpurun_out/r2_gpu_tests.log | cut -c1-200
tail -n 20 gpurun_out/r2_sweep.log
