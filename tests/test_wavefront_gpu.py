"""GPU, EXPERIMENTAL (round-2 candidate): the wavefront schedule of the in-order epoch
(fm_sgd_inorder_wavefront_kernel, fmb200_set_tuning variant 4) against the sequential
oracle, bit for bit.  The schedule itself is proven equivalent on the CPU
(tests/test_oracle.py::test_wavefront_schedule_is_sequentially_equivalent); the kernel was
written after this round's GPU budget was spent and has not run on a device yet, so it is
opt-in in the library and this module only runs with FMB200_EXPERIMENTAL=1:

    FMB200_EXPERIMENTAL=1 python -m pytest tests/test_wavefront_gpu.py -q
"""
import os
import time

import numpy as np
import pytest

from conftest import make_learner
from libfm_b200 import MODE_INORDER, synth
from oracle import Port
from test_oracle import _ragged_short_rows

pytestmark = [pytest.mark.gpu]  # validated on the device in round 2: default for eligible shapes


def _case(case):
    task, k, k0, k1, regs = 0, 8, 1, 1, (0.0, 0.0, 0.0)
    if case == "c2_shape":
        tr = synth.two_field(200_000, 6040, 3706, seed=3)
    elif case == "zipf":
        tr = synth.two_field(40_000, 600, 400, seed=4, zipf=1.1)
    elif case == "ragged":
        tr = _ragged_short_rows(30_000, 500, seed=5)
    elif case == "dups":
        tr = _ragged_short_rows(30_000, 300, seed=6, dup_every=7)
    elif case == "classification":
        tr = synth.two_field(30_000, 800, 500, seed=7)
        tr.target[:] = np.where(tr.target > 3, 1.0, -1.0)
        task = 1
    elif case == "no_bias":
        tr = synth.two_field(30_000, 800, 500, seed=8)
        k0, k1 = 0, 0
    elif case == "tiny":
        tr = synth.two_field(5, 3, 3, seed=1)
    else:
        tr = _ragged_short_rows(20_000, 400, seed=9)
        k, regs = 3, (0.01, 0.02, 0.03)
    return tr, task, k, k0, k1, regs


@pytest.mark.parametrize("case", ["tiny", "c2_shape", "zipf", "ragged", "dups", "classification", "no_bias", "k3_reg"])
def test_wavefront_kernel_bit_exact(case, built_lib):
    tr, task, k, k0, k1, regs = _case(case)
    n = tr.num_feature
    r = np.random.default_rng(1)
    init = (0.05, r.standard_normal(n) * 0.1, r.standard_normal((k, n)) * 0.1)
    mn, mx = float(tr.target.min()), float(tr.target.max())
    cfg = dict(n=n, k=k, k0=k0, k1=k1, task=task, lr=0.02, regs=np.array(regs), min_target=mn, max_target=mx)
    p = Port(n, k, k0, k1)
    p.set_params(*init)
    p.reg0, p.regw, p.regv = regs
    l = make_learner(cfg, init, mode=MODE_INORDER)
    l.set_tuning(variant=4)
    for _ in range(2):
        p.sgd_epoch(tr, task, 0.02, mn, mx)
        l.sgd_epoch(tr)
        assert l.epoch_config()["slots"] == 4 and l.epoch_config()["rows_per_tile"] == 32  # the wavefront kernel ran
    l.pull_params()
    if task == 0:
        assert l.fm.w0 == p.w0.value
        assert np.array_equal(l.fm.w, p.w) and np.array_equal(l.fm.v, p.v)
    else:  # device exp() vs glibc exp()
        np.testing.assert_allclose(l.fm.w, p.w, rtol=0, atol=1e-12)
        np.testing.assert_allclose(l.fm.v, p.v, rtol=0, atol=1e-12)
    l.close()


def test_wavefront_kernel_speed_report(built_lib):
    """Not an assertion of speed: prints both in-order kernels' epoch time on the C2 shape."""
    tr = synth.movielens_1m_shaped()
    n, k = tr.num_feature, 8
    r = np.random.default_rng(1)
    init = (0.0, np.zeros(n), r.standard_normal((k, n)) * 0.1)
    cfg = dict(n=n, k=k, k0=1, k1=1, task=0, lr=0.01, regs=np.zeros(3), min_target=1.0, max_target=5.0)
    out = {}
    for variant in (1, 4):
        l = make_learner(cfg, init, mode=MODE_INORDER)
        l.set_tuning(variant=variant)
        l.sgd_epoch(tr)
        t = time.perf_counter()
        l.sgd_epoch(tr)
        out[variant] = time.perf_counter() - t
        l.pull_params()
        out[("w0", variant)] = l.fm.w0
        l.close()
    print("\nin-order C2 epoch: row-at-a-time %.3f s, wavefront %.3f s" % (out[1], out[4]))
    assert out[("w0", 1)] == out[("w0", 4)]
