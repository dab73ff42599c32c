#!/bin/bash
# Round-2 GPU call on 8 GPUs (second): the final exchange (local mean-count table, parity-buffered counts) at N = 8 and 4.
set -u
mkdir -p gpurun_out
run() { n=$1; port=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --steps 30 --warmup 5 "$@"; }
run 8 29541 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err; echo "n8 rc=$?"; tail -n 2 gpurun_out/r2_bench_n8b.err
timeout 300 python -m pytest tests/test_hogwild_gpu.py tests/test_cli_gpu.py -m gpu -q -k "peer or two_gpus" 2>&1 | tail -n 2
python - <<'PY'
import json
for f in ("n8b",):
    try:
        d = json.loads(open("gpurun_out/r2_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f e2e %.4g launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
        pm = d.get("parity_multi_gpu")
        if pm: print("  heldout gpu", ["%.4f" % x for x in pm["heldout_rmse_gpu"]], "seq", ["%.4f" % x for x in pm["heldout_rmse_one_sequential_stream"]])
        c5 = (d.get("extra") or {}).get("c5")
        if c5: print("  c5:", {k: v for k, v in c5.items() if k in ("value", "ms_per_step", "ms_epoch_kernel", "ms_exchange", "error")})
    except Exception as e:
        print(f, "unreadable:", e)
PY
