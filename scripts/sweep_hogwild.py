"""On-GPU sweep of the HOGWILD epoch kernel on BASELINE config C2 (planted-signal ratings, so there is
something to learn): for each launch geometry / kernel variant, the epoch time and the RMSE gap to the
sequential oracle's trajectory from the same initial model.  The default geometry is chosen from THIS
curve, not from speed alone.

    python scripts/sweep_hogwild.py [--epochs 6] [--out gpurun_out/r2_sweep.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from libfm_b200 import MODE_HOGWILD, FmLearnSgdElement, FmModel, synth  # noqa: E402
from oracle import Port  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--out", default="gpurun_out/r2_sweep.json")
    ap.add_argument("--zipf", type=float, default=0.0)
    args = ap.parse_args()
    tr, te = synth.movielens_1m_planted(100_000, seed=7, zipf=args.zipf)
    n, k = tr.num_feature, 8
    v0 = np.random.default_rng(42).standard_normal((k, n)) * 0.1
    port = Port(n, k)
    port.set_params(0.0, np.zeros(n), v0)
    ref = []
    for _ in range(args.epochs):
        port.sgd_epoch(tr, 0, 0.01, tr.min_target, tr.max_target)
        ref.append((port.metric(tr, 0, tr.min_target, tr.max_target),
                    port.metric(te, 0, tr.min_target, tr.max_target)))
    print("oracle trajectory (train, test):", ["%.5f/%.5f" % r for r in ref], flush=True)

    # (ctas_per_sm, rows_per_tile, threads, damp, variant)
    tunings = [(0, 0, 0, 0, 0), (0, 0, 0, 0, 3), (0, 0, 128, 0, 0), (0, 0, 64, 0, 0), (2, 0, 0, 0, 0), (1, 0, 0, 0, 0),
               (2, 0, 128, 0, 0), (1, 0, 128, 0, 0), (1, 0, 64, 0, 0), (1, 0, 32, 0, 0)]
    rows = []
    for t in tunings:
        fm = FmModel(n, k)
        fm.v = v0.copy()
        l = FmLearnSgdElement(fm, mode=MODE_HOGWILD)
        l.task, l.learn_rate = 0, 0.01
        l.min_target, l.max_target = tr.min_target, tr.max_target
        l.push_hparams()
        l.set_tuning(*t)
        secs, gaps, traj = [], [], []
        for e in range(args.epochs):
            secs.append(l.sgd_epoch(tr))
            g = (l.evaluate(tr), l.evaluate(te))
            traj.append(g)
            gaps.append(max(abs(g[0] - ref[e][0]), abs(g[1] - ref[e][1])))
        cfg = l.epoch_config()
        window = min(tr.num_cases, cfg["grid"] * cfg["rows_per_tile"])
        row = {"tuning": t, "cfg": cfg, "window_rows": window, "best_us": min(secs[1:]) * 1e6,
               "gex_s": tr.num_cases / min(secs[1:]) / 1e9, "gap_first": gaps[0], "gap_last": gaps[-1],
               "gap_max": max(gaps), "rmse_last": traj[-1]}
        rows.append(row)
        print("tune=%-22s grid=%4d x %3d window=%7d  %7.1f us %6.2f Gex/s  gap e0 %.4f  e%d %.4f  max %.4f" % (
            t, cfg["grid"], cfg["block"], window, row["best_us"], row["gex_s"], gaps[0], args.epochs - 1,
            gaps[-1], max(gaps)), flush=True)
        l.close()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump({"oracle": ref, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
