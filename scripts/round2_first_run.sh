#!/bin/bash
# First GPU call of the next round, as ONE gpurun command (one box acquisition):
#   gpurun --timeout 900 -- 'bash scripts/round2_first_run.sh'
# 1. the staged wavefront in-order kernel (opt-in variant 4): bit-exact parity + timing report
# 2. the regular GPU suite (must stay green whatever 1. says)
# 3. a short bench line
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
FMB200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_wavefront_gpu.py -q -s > gpurun_out/r2_wavefront.log 2>&1
echo "wavefront rc=$?" | tee -a gpurun_out/r2_wavefront.log
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 200 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -3 gpurun_out/r2_wavefront.log gpurun_out/r2_gpu_tests.log gpurun_out/r2_bench.json
