#!/bin/bash
# Round-2 GPU call A: latency microbench, ORDERED-mode parity + timing, upload paths, staged wavefront,
# full suite, hogwild variant sweep, bench.
set -u
mkdir -p gpurun_out
timeout 60 ./scripts/micro/lat_bench > gpurun_out/r2_lat_bench.txt 2>&1
timeout 900 python -m pytest tests/test_ordered_gpu.py -q -s -x > gpurun_out/r2_ordered.log 2>&1
echo "ordered rc=$?" | tee -a gpurun_out/r2_ordered.log
timeout 400 python -m pytest tests/test_upload_gpu.py tests/test_mcmc_gpu.py -q -s > gpurun_out/r2_upload.log 2>&1
echo "upload rc=$?" | tee -a gpurun_out/r2_upload.log
FMB200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_wavefront_gpu.py -q -s > gpurun_out/r2_wavefront.log 2>&1
echo "wavefront rc=$?" | tee -a gpurun_out/r2_wavefront.log
timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_ordered_gpu.py --deselect tests/test_upload_gpu.py --deselect tests/test_mcmc_gpu.py > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 400 python scripts/sweep_hogwild.py --out gpurun_out/r2_sweep.json > gpurun_out/r2_sweep.log 2>&1
echo "sweep rc=$?" | tee -a gpurun_out/r2_sweep.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r2_smoke.log
tail -4 gpurun_out/r2_ordered.log gpurun_out/r2_upload.log gpurun_out/r2_wavefront.log gpurun_out/r2_gpu_tests.log gpurun_out/r2_bench.err
grep "\[ordered" gpurun_out/r2_ordered.log
cat gpurun_out/r2_sweep.log | tail -20
