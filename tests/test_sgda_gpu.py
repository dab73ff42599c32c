"""GPU: SGDA (reference fm_learn_sgd_element_adapt_reg.h) through
fmb200_sgda_begin / _epoch / _get_reg against the oracle restatement (oracle/fm_oracle_sgda.c, pinned
bit-identical to the reference's own learner by tests/test_oracle.py).  Bar: bit-exact parameters and
regularisation values for regression; classification to 1e-12 (device exp())."""
import numpy as np
import pytest

from conftest import make_learner
from libfm_b200 import MODE_INORDER, synth
from oracle import Port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["two_groups_reg", "one_group", "classification", "k40_ragged", "no_bias_no_linear"])
def test_sgda_matches_oracle(case, built_lib):
    task, k, k0, k1, groups = 0, 5, 1, 1, 2
    if case == "k40_ragged":
        full = synth.ragged(5000, 300, 6, seed=14)
        k, groups = 40, 3
    else:
        full = synth.two_field(12_000, 300, 200, seed=4, planted_k=3)
    tr, rest = synth.split_rows(full, full.num_cases * 2 // 3)
    va, te = synth.split_rows(rest, rest.num_cases // 2)   # validation shorter than train: the cursor wraps
    if case == "one_group":
        groups = 1
    if case == "classification":
        task = 1
        for d in (tr, va):
            d.target[:] = np.where(d.target > 3, 1.0, -1.0)
    if case == "no_bias_no_linear":
        k0, k1 = 0, 0
    n = full.num_feature
    group = (np.arange(n) * groups // n).astype(np.uint32)
    mn, mx = float(tr.target.min()), float(tr.target.max())
    r = np.random.default_rng(2)
    init = (0.0, np.zeros(n), r.standard_normal((k, n)) * 0.1)
    cfg = dict(n=n, k=k, k0=k0, k1=k1, task=task, lr=0.02, regs=np.zeros(3), min_target=mn, max_target=mx)
    p = Port(n, k, k0, k1)
    p.set_params(*init)
    p.sgda_begin(group)
    l = make_learner(cfg, init, mode=MODE_INORDER)
    l.sgda_begin(group if groups > 1 else None)
    for e in range(3):
        sec = l.sgda_epoch(tr, va, e > 0)
        p.sgda_epoch(tr, va, task, 0.02, mn, mx, e > 0)
    l.pull_params()
    reg_w, reg_v = l.sgda_reg()
    if task == 0:
        assert l.fm.w0 == p.w0.value and np.array_equal(l.fm.w, p.w) and np.array_equal(l.fm.v, p.v)
        assert np.array_equal(reg_w, p.reg_w) and np.array_equal(reg_v, p.reg_v)
    else:
        np.testing.assert_allclose(l.fm.v, p.v, rtol=0, atol=1e-12)
        np.testing.assert_allclose(reg_v, p.reg_v, rtol=0, atol=1e-12)
        np.testing.assert_allclose(reg_w, p.reg_w, rtol=0, atol=1e-12)
    assert (reg_v.max() > 0) or k == 0
    print("\n[sgda %s] %d train rows: %.2f ms/epoch" % (case, tr.num_cases, sec * 1e3))
    l.close()
