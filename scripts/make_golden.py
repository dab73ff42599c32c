"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libfm_ref.so,
i.e. /root/reference compiled in place).  Run in the authoring container only:

    python scripts/make_golden.py

Each fixture stores the seeded input (CSR), the hyper-parameters and what the
reference's fm_learn_sgd_element::learn / evaluate / predict produced: initial
and final w0/w/V (factor-major), per-epoch train/test metric in full double
precision, and the final transformed predictions on the test set.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libfm_b200 import synth  # noqa: E402
from oracle import Ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (train, test, task, k, k0, k1, lr, regs, init_stdev, seed, epochs)
    "reg_k8_twofield": (synth.two_field(600, 40, 30, 1), synth.two_field(150, 40, 30, 2),
                        0, 8, 1, 1, 0.01, (0.0, 0.0, 0.0), 0.1, 42, 3),
    "reg_k8_regularised": (synth.two_field(600, 40, 30, 3, planted_k=4), synth.two_field(150, 40, 30, 4, planted_k=4),
                           0, 8, 1, 1, 0.02, (0.01, 0.02, 0.03), 0.1, 7, 3),
    "reg_k5_ragged": (synth.ragged(400, 64, 9, 5), synth.ragged(100, 64, 9, 6),
                      0, 5, 1, 1, 0.005, (0.0, 0.0, 0.01), 0.1, 11, 2),
    "reg_k16_nobias_nolinear": (synth.ragged(300, 50, 6, 8, empty_frac=0.0), synth.ragged(80, 50, 6, 9),
                                0, 16, 0, 0, 0.01, (0.0, 0.0, 0.0), 0.05, 3, 2),
    "cls_k4_multifield": (synth.multi_field(500, 5, 100, 12), synth.multi_field(120, 5, 100, 13),
                          1, 4, 1, 1, 0.05, (0.0, 0.01, 0.01), 0.1, 5, 3),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (tr, te, task, k, k0, k1, lr, regs, stdev, seed, epochs) in CASES.items():
        if task == 1:
            tr.binarize_targets()
            te.binarize_targets()
        n = max(tr.num_feature, te.num_feature)  # libfm.cpp:203
        ref = Ref(n, k, k0=k0, k1=k1, init_stdev=stdev, seed=seed)
        ref.set_reg(*regs)
        w0_0, w_0, v_0 = ref.get_params()
        mn, mx = tr.min_target, tr.max_target  # libfm.cpp:295-296
        hist_tr, hist_te = [], []
        for _ in range(epochs):
            ref.learn(tr, te, task, lr, 1, mn, mx)
            hist_tr.append(ref.evaluate(tr, task, mn, mx))
            hist_te.append(ref.evaluate(te, task, mn, mx))
        w0, w, v = ref.get_params()
        pred = ref.predict(te, task, mn, mx)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            tr_row_ptr=tr.row_ptr, tr_col=tr.col, tr_val=tr.val, tr_target=tr.target,
            te_row_ptr=te.row_ptr, te_col=te.col, te_val=te.val, te_target=te.target,
            n=n, k=k, k0=k0, k1=k1, task=task, lr=lr, regs=np.array(regs), init_stdev=stdev,
            seed=seed, epochs=epochs, min_target=mn, max_target=mx,
            w0_init=w0_0, w_init=w_0, v_init=v_0, w0=w0, w=w, v=v,
            metric_train=np.array(hist_tr), metric_test=np.array(hist_te), pred_test=pred)
        print(name, "train", hist_tr[-1], "test", hist_te[-1])


if __name__ == "__main__":
    main()
