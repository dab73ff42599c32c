#!/bin/bash
# Round-2 GPU call on 2 GPUs (third, short): the exchange with the rows mean in the first kernel: peer tests + bench N = 2.
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_hogwild_gpu.py tests/test_cli_gpu.py -m gpu -q -k "peer or two_gpus" 2>&1 | tail -n 3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err
echo "bench n2 rc=$?"; tail -n 2 gpurun_out/r2_bench_n2c.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2c.json").read().strip().splitlines()[-1])
print("value %.4g ms %.4f e2e %.4g launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
print("  parity_multi_gpu:", d["parity_multi_gpu"]["heldout_rmse_gpu"], d["parity_multi_gpu"]["max_abs_gap"])
PY
