"""CPU study (oracle only, no GPU): how should G row-sharded replicas be combined once per epoch?

Each replica runs the reference's sequential SGD (oracle/fm_oracle.c) on its shard from the common
state theta0; the exchange then forms  theta = theta0 + gamma * sum_g (theta_g - theta0)  per parameter.

  average     gamma = 1/G                      (what fmb200_allreduce_mean does today)
  sum         gamma = 1                        (Hogwild across GPUs; overshoots saturated parameters)
  saturation  gamma_i = 1 / (1 + (G-1) s_i),   s_i = 1 - exp(-lr * h_i * count_i)
  meanfield   gamma_i = (1 - (1-s_i)^G) / (G s_i), same s_i
              count_i = occurrences of feature i in the shard (the per-slot table the library already
              builds for its in-GPU damping), h_i = 1 for w0 / w (x = 1), mean |v|^2 for V rows.
              A parameter whose shard-epoch has already converged it (s -> 1: the bias, hot features)
              is averaged; one that was barely touched (s -> 0) is summed.

Prints test RMSE per epoch against the single-stream sequential run (the reference's trajectory).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from libfm_b200 import synth
from libfm_b200.model import Data
from oracle import Port


def shard(d, g, G):
    lo, hi = d.num_cases * g // G, d.num_cases * (g + 1) // G
    b, e = int(d.row_ptr[lo]), int(d.row_ptr[hi])
    return Data((d.row_ptr[lo:hi + 1] - d.row_ptr[lo]).astype(np.uint64), d.col[b:e].copy(), d.val[b:e].copy(),
                d.target[lo:hi].copy(), d.num_feature)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=400_000)
    ap.add_argument("--users", type=int, default=3000)
    ap.add_argument("--items", type=int, default=2000)
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=12)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--zipf", type=float, default=0.0)
    a = ap.parse_args()
    full = synth.two_field(a.rows + a.rows // 5, a.users, a.items, seed=5, planted_k=4, zipf=a.zipf)
    tr, te = shard(full, 0, 1), None
    n_tr = a.rows
    b = int(full.row_ptr[n_tr])
    tr = Data(full.row_ptr[:n_tr + 1].copy(), full.col[:b].copy(), full.val[:b].copy(), full.target[:n_tr].copy(),
              full.num_feature)
    te = Data((full.row_ptr[n_tr:] - full.row_ptr[n_tr]).astype(np.uint64), full.col[b:].copy(), full.val[b:].copy(),
              full.target[n_tr:].copy(), full.num_feature)
    n, k, G = tr.num_feature, 8, a.gpus
    shards = [shard(tr, g, G) for g in range(G)]
    counts = [np.bincount(s.col, minlength=n).astype(np.float64) for s in shards]
    mn, mx = float(tr.target.min()), float(tr.target.max())

    def fresh():
        p = Port(n, k)
        p.init(42, 0.0, 0.1)
        return p

    seq = fresh()
    rules = ["average", "sum", "saturation", "meanfield"]
    state = {r: fresh() for r in rules}
    print("epoch  sequential " + " ".join("%11s" % r for r in rules))
    for e in range(a.epochs):
        seq.sgd_epoch(tr, 0, a.lr, mn, mx)
        row = [seq.metric(te, 0, mn, mx)]
        for r in rules:
            p0 = state[r]
            w0_0, w_0, v_0 = p0.w0.value, p0.w.copy(), p0.v.copy()
            d_w0, d_w, d_v = 0.0, np.zeros_like(w_0), np.zeros_like(v_0)
            for g in range(G):
                q = Port(n, k)
                q.set_params(w0_0, w_0, v_0)
                q.sgd_epoch(shards[g], 0, a.lr, mn, mx)
                d_w0 += q.w0.value - w0_0
                d_w += q.w - w_0
                d_v += q.v - v_0
            if r == "average":
                g0 = gw = gv = 1.0 / G
            elif r == "sum":
                g0 = gw = gv = 1.0
            else:
                # meanfield: the factor that makes G summed shard-steps of relative size s equal G such
                # steps taken one after the other on a quadratic: (1 - (1-s)^G) / (G s) -- the same
                # closed form the HOGWILD kernels use inside one GPU
                cnt = np.mean(counts, axis=0)
                sat = lambda h, c: 1.0 - np.exp(-a.lr * h * c)  # noqa: E731
                if r == "saturation":
                    gam = lambda s_: 1.0 / (1.0 + (G - 1) * s_)  # noqa: E731
                else:
                    gam = lambda s_: np.where(s_ > 1e-9, (1.0 - (1.0 - s_) ** G) / (G * np.maximum(s_, 1e-9)), 1.0)  # noqa: E731
                g0 = float(gam(np.float64(sat(1.0, tr.num_cases / G))))
                gw = gam(sat(1.0, cnt))
                hv = float(np.mean(np.sum(v_0 * v_0, axis=0)))  # mean squared norm of a factor row
                gv = gam(sat(hv, cnt))[None, :]
            p0.set_params(w0_0 + g0 * d_w0, w_0 + gw * d_w, v_0 + gv * d_v)
            row.append(p0.metric(te, 0, mn, mx))
        print("%5d  %10.5f " % (e + 1, row[0]) + " ".join("%11.5f" % x for x in row[1:]), flush=True)


if __name__ == "__main__":
    main()
