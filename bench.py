#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the libFM SGD hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], the configuration `metric` is quoted on):
SGD, k=8, MovieLens-1M-shaped CSR (6040 users x 3706 items, 1,000,209 rows,
2 nnz/row, value 1), regression, lr 0.01, init_stdev 0.1.  A "step" is one pass
of the hot path over the whole data set = one SGD epoch = one launch of
fm_sgd_hogwild_kernel (plus, for N > 1, the per-epoch NCCL all-reduce of w0|w|V
and the 1/N scale).  Weak scaling: every rank owns a full C2-sized row shard.

value  : examples/s with inputs resident in HBM, CUDA events on the library's own
         stream, L2 flushed (256 MiB write) before every timed step, max over ranks.
e2e    : examples/s through the C ABI with HOST (pinned) buffers: every step uploads
         the CSR (fmb200_upload_data_async, two device slots so the copy of the next
         step overlaps this step's epoch), runs the epoch and reads the model back
         (fmb200_get_params); wall clock over the timed steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MovieLens-1M-shaped examples/sec at k=8, SGD"
UNIT = "examples/s"
K_FACTORS = 8
LEARN_RATE = 0.01
WORKLOAD = {
    "workload": "C2: libFM SGD epoch, k=8, MovieLens-1M-shaped CSR (6040 users x 3706 items, "
                "1000209 rows, 2 nnz/row, x=1), -task r -learn_rate 0.01 -init_stdev 0.1",
    "rows_per_gpu": 1_000_209, "k": K_FACTORS, "nnz_per_row": 2, "mode": "hogwild",
    "l2": "flushed before every timed step (256 MiB write)",
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation
# --------------------------------------------------------------------------
def cpu_reference_epochs(data, n_epochs):
    """Time the row loop of fm_learn_sgd_element::learn on the host.

    Uses oracle/_ref (the unmodified reference compiled in place) when present,
    else the C restatement.  Returns (examples_per_sec, kind, per_epoch_seconds).
    The reference is single-threaded, so cores == 1 is all it can use."""
    import numpy as np
    import oracle
    n = data.num_feature
    if oracle.have_ref():
        ref = oracle.Ref(n, K_FACTORS, seed=42, init_stdev=0.1)
        # time_learn = the reference's own user-CPU clock around the row loop
        # (fm_learn_sgd_element.h:55,68); evaluate passes are outside it
        tiny = data.rows(0, 1)
        _, _, tm = ref.learn(data, tiny, 0, LEARN_RATE, n_epochs, data.min_target, data.max_target)
        secs = [float(t) for t in tm]
        kind = "reference"
    else:
        port = oracle.Port(n, K_FACTORS)
        port.init(42, 0.0, 0.1)
        secs = []
        for _ in range(n_epochs):
            t0 = time.process_time()
            port.sgd_epoch(data, 0, LEARN_RATE, data.min_target, data.max_target)
            secs.append(time.process_time() - t0)
        kind = "port"
    best = statistics.median(secs)
    return data.num_cases / best, kind, secs


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0  # the reference has no multi-process path: rank 0 alone measures
    from libfm_b200 import synth
    data = synth.movielens_1m_shaped(seed=7)
    n_ep = args.warmup + args.steps
    t0 = time.time()
    ex_s, kind, secs = cpu_reference_epochs(data, n_ep)
    timed = secs[args.warmup:]
    ms = 1e3 * sum(timed) / len(timed)
    value = data.num_cases / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(WORKLOAD, mode="reference CPU, in-order"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": "%d full epochs of the 1000209-row workload, the reference's own "
                                   "time_learn (user CPU s around the row loop)" % len(timed)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------
# clocks: NVML sampling thread
# --------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if not self.nv:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for b, name in self.REASONS.items():
                if bits & b:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.01)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


class _DevBuf:
    """Expose a raw device pointer to torch through the CUDA array interface."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4",
                                         "data": (ptr, False), "version": 2}


# --------------------------------------------------------------------------
# the GPU arm
# --------------------------------------------------------------------------
def run_gpu_arm(args):
    import numpy as np
    import torch
    from libfm_b200 import FmLearnSgdElement, FmModel, MODE_HOGWILD, synth

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # every rank owns its own full-size shard (weak scaling), seeded by rank
    data = synth.movielens_1m_shaped(seed=7 + rank)
    n = data.num_feature
    fm = FmModel(n, K_FACTORS)
    fm.init_stdev = 0.1
    fm.init_numpy(42)  # identical replicas on every rank
    lrn = FmLearnSgdElement(fm, device=local_rank, mode=MODE_HOGWILD)
    lrn.task, lrn.learn_rate = 0, LEARN_RATE
    lrn.min_target, lrn.max_target = data.min_target, data.max_target
    lrn.push_hparams()
    lrn.upload(data, 0)

    stream = torch.cuda.ExternalStream(lrn.stream(), device=local_rank)
    ptr, n_floats = lrn.params_device()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    lib, ctx = lrn.lib, lrn._ctx
    import ctypes as C

    # the per-epoch exchange: one-shot all-reduce over NVLink peer memory (fm_peer.cu),
    # NCCL through torch.distributed as the fallback / comparison
    collective = "none"
    params = None
    if world > 1:
        collective = args.collective
        if collective in ("auto", "p2p"):
            h = C.create_string_buffer(64)
            ok = lib.fmb200_peer_export(ctx, h) == 0
            handles = [None] * world
            dist.all_gather_object(handles, h.raw if ok else b"")
            ok = ok and all(len(x) == 64 for x in handles)
            if ok:
                ok = lib.fmb200_peer_attach_ipc(ctx, world, rank, b"".join(handles)) == 0
            flag = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                collective = "p2p"
            elif collective == "p2p":
                raise RuntimeError("peer attach failed: " + lib.fmb200_last_error().decode())
            else:
                collective = "nccl"
        if collective == "nccl":
            params = torch.as_tensor(_DevBuf(ptr, n_floats), device=torch.device("cuda", local_rank))

    def step():
        rc = lib.fmb200_sgd_epoch_async(ctx, 0)
        if rc == 0 and collective == "p2p":
            rc = lib.fmb200_allreduce_mean(ctx)
        elif rc == 0 and collective == "nccl":
            dist.all_reduce(params)  # sum over ranks on the library's stream
            rc = lib.fmb200_scale_params(ctx, 1.0 / world)
        if rc != 0:
            raise RuntimeError(lib.fmb200_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tick = torch.zeros(1, device="cuda")

    def align():
        # the L2 flush takes a different time on every rank; without re-aligning, the
        # first rank to start its epoch would be charged the other ranks' flush time
        # while it waits in the exchange
        if collective == "p2p":
            if lib.fmb200_peer_barrier(ctx) != 0:
                raise RuntimeError(lib.fmb200_last_error().decode())
        elif collective == "nccl":
            dist.all_reduce(tick)

    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            flush.zero_()
            step()
        barrier()
        launches0 = lrn.kernel_launches()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
        sampler.start()
        wall0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()          # evict the CSR and the parameters from L2 (untimed)
            align()                # N > 1: all ranks enter the step together (untimed)
            a.record(stream)
            step()
            b.record(stream)
        torch.cuda.synchronize()
        sampler.sample()
        sampler.stop_flag = True
        wall = time.perf_counter() - wall0
        barrier()
    launches = lrn.kernel_launches() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    rows = data.num_cases
    value = world * rows / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI, host buffers ------------------------
    from libfm_b200.model import pinned_copy
    rp, col, val, tgt = [pinned_copy(a) for a in (data.row_ptr, data.col, data.val, data.target)]
    w0 = C.c_double()
    w_out = np.empty(n, dtype=np.float64)
    v_out = np.empty((K_FACTORS, n), dtype=np.float64)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    n_e2e = max(3, min(args.steps, 20))

    def upload_async(slot):
        rc = lib.fmb200_upload_data_async(ctx, slot, rows, int(rp[-1]), P(rp, C.c_uint64), P(col, C.c_uint32),
                                          P(val, C.c_float), P(tgt, C.c_float))
        if rc != 0:
            raise RuntimeError(lib.fmb200_last_error().decode())

    # Two device slots, ping-pong: while the epoch of step i runs on slot A, the inputs of
    # step i+1 are already crossing PCIe into slot B (copy stream).  Every step still
    # pays one full upload, one epoch (+ exchange) and one read-back of the model.
    cur_slot = [2, 3]

    def e2e_step():
        upload_async(cur_slot[1])                 # next step's inputs: host -> device
        global_slot = cur_slot[0]
        rc = lib.fmb200_sgd_epoch_async(ctx, global_slot)   # waits for this slot's upload
        if rc == 0 and collective == "p2p":
            rc = lib.fmb200_allreduce_mean(ctx)
        elif rc == 0 and collective == "nccl":
            with torch.cuda.stream(stream):
                dist.all_reduce(params)
            rc = lib.fmb200_scale_params(ctx, 1.0 / world)
        if rc == 0:
            rc = lib.fmb200_get_params(ctx, C.byref(w0), P(w_out, C.c_double), P(v_out, C.c_double))
        if rc != 0:
            raise RuntimeError(lib.fmb200_last_error().decode())
        cur_slot.reverse()

    upload_async(cur_slot[0])
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    barrier()
    e2e_value = world * rows / float(e2e_s.item())
    h2d = int(rp.nbytes + col.nbytes + val.nbytes + tgt.nbytes)
    d2h = int(n_floats * 4)

    cfg = lrn.epoch_config()
    lrn.close()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (fm_sgd_hogwild_kernel) -------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_example = 2 * K_FACTORS * 2 * 4  # 2*k*nnz*4
    if world == 1:
        kernel_ms = ms_per_step  # the step is exactly one launch of the epoch kernel
    else:
        kernel_ms = ms_per_step  # epoch kernel + all-reduce + scale; the kernel dominates
    achieved = rows * bytes_per_example / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "c2_epoch_dram_bytes.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_example": bytes_per_example,
                "kernel": "fm_sgd_hogwild_kernel<G=%d,S=%d>" % (cfg["lanes_per_row"], cfg["slots"])}

    # ---- CPU baseline on this box's host cores (N == 1 only) ------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = 40  # ~75 ms of row loop each plus the reference's own evaluate passes: ~10-20 s
        ex_s, kind, secs = cpu_reference_epochs(synth.movielens_1m_shaped(seed=7), n_cpu)
        cpu = {"value": ex_s, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "%d full epochs of the 1000209-row C2 workload (median time_learn %.1f ms)"
                         % (n_cpu, 1e3 * statistics.median(secs))}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(WORKLOAD, parallelism="row-sharded dp%d, per-epoch all-reduce of w0|w|V (%s)" % (
            world, {"p2p": "one-shot kernel over NVLink peer memory", "nccl": "NCCL", "none": "single GPU"}[collective]),
                       kernel_geometry=cfg),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n_e2e},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "timed_region_wall_s": wall,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="auto", choices=["auto", "p2p", "nccl"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
