#!/bin/bash
# Round-2 GPU call on 2 GPUs (second): the exchange with the local mean-count table: peer tests, bench N = 2.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_hogwild_gpu.py tests/test_cli_gpu.py -m gpu -q -s -k "peer or two_gpus" > gpurun_out/r2_tests_n2b.log 2>&1
echo "tests rc=$?"; tail -n 4 gpurun_out/r2_tests_n2b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 30 --warmup 5 --c5 --c5-rows 25000000 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
echo "bench n2 rc=$?"; tail -n 2 gpurun_out/r2_bench_n2b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2b.json").read().strip().splitlines()[-1])
print("value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"]["parallelism"], "launches", d["gpu_launches"])
print("  parity_multi_gpu:", d["parity_multi_gpu"]["heldout_rmse_gpu"], d["parity_multi_gpu"]["max_abs_gap"])
c5 = d["extra"]["c5"]; print("  c5:", {k: v for k, v in c5.items() if k in ("value", "ms_per_step", "ms_epoch_kernel", "ms_exchange", "error")})
PY
