#!/bin/bash
# Round-2 GPU call on 8 GPUs: bench.py under torchrun at N = 8 (with BASELINE config C5) and N = 4 back to back
# on ONE box, the NCCL exchange for comparison at N = 8 (N = 2 has its own 2-GPU call: 8-GPU minutes cost 8x).
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r2_topo.txt
run() {  # n port extra...
  n=$1; port=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --steps 30 --warmup 5 "$@"
}
run 8 29521 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "n8 rc=$?"; tail -n 2 gpurun_out/r2_bench_n8.err
run 4 29522 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "n4 rc=$?"
run 8 29524 --collective nccl --no-parity --no-extras > gpurun_out/r2_bench_n8_nccl.json 2> gpurun_out/r2_bench_n8_nccl.err; echo "n8 nccl rc=$?"
timeout 300 python -m pytest tests/test_cli_gpu.py -m gpu -q -s -k "two_gpus" 2>&1 | tail -n 4
python - <<'PY'
import json
for f in ("n8", "n4", "n8_nccl"):
    try:
        d = json.loads(open("gpurun_out/r2_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"]["parallelism"])
        pm = d.get("parity_multi_gpu")
        if pm: print("  heldout gpu", ["%.4f" % x for x in pm["heldout_rmse_gpu"]], "seq", ["%.4f" % x for x in pm["heldout_rmse_one_sequential_stream"]])
        c5 = (d.get("extra") or {}).get("c5")
        if c5: print("  c5:", {k: v for k, v in c5.items() if k not in ("kernel_geometry", "workload", "roofline")}, c5.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
