#!/bin/bash
# Round-2 GPU call on 2 GPUs: bench.py under torchrun (N = 2; C5 shape at 12.5 M rows per GPU), NCCL comparison,
# the 2-GPU CLI test, the two-context peer tests.
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 60 python scripts/prof_ordered.py 200000 0 132 2>&1 | cut -c1-420 | tee gpurun_out/r2_ordered_v11.txt
timeout 60 python scripts/prof_ordered.py 1000209 0 0 2>&1 | cut -c1-120 | tee -a gpurun_out/r2_ordered_v11.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 30 --warmup 5 --c5 --c5-rows 25000000 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?"; tail -n 3 gpurun_out/r2_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 30 --warmup 5 --collective nccl --no-parity > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err
echo "bench n2 nccl rc=$?"; tail -n 3 gpurun_out/r2_bench_n2_nccl.err
timeout 600 python -m pytest tests/test_cli_gpu.py tests/test_hogwild_gpu.py -m gpu -q -s -k "two_gpus or peer" > gpurun_out/r2_tests_n2.log 2>&1
echo "tests rc=$?"; tail -n 6 gpurun_out/r2_tests_n2.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n2.json", "gpurun_out/r2_bench_n2_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"]["parallelism"])
        print("  parity_multi_gpu:", d.get("parity_multi_gpu"))
        print("  c5:", {k: v for k, v in (d.get("extra", {}).get("c5") or {}).items() if k not in ("kernel_geometry", "workload")})
    except Exception as e:
        print(f, "unreadable:", e)
PY
