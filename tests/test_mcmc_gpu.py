"""GPU: the MCMC / ALS e-term pass (reference fm_learn_mcmc.h:148-378) through
fmb200_mcmc_eterms, against the oracle restatement (oracle/fm_oracle.c::fmo_mcmc_eterms, pinned
bit-identical to the reference's own pass by tests/test_oracle.py).  Bar: bit-exact."""
import time

import numpy as np
import pytest

from conftest import make_learner
from libfm_b200 import MODE_HOGWILD, MODE_INORDER, FmError, synth
from oracle import Port

pytestmark = pytest.mark.gpu


def _setup(d, k, k0=1, k1=1, seed=1):
    n = d.num_feature
    r = np.random.default_rng(seed)
    init = (0.3, r.standard_normal(n) * 0.1, r.standard_normal((k, n)) * 0.1)
    cfg = dict(n=n, k=k, k0=k0, k1=k1, task=0, lr=0.01, regs=np.zeros(3), min_target=1.0, max_target=5.0)
    p = Port(n, k, k0, k1)
    p.set_params(*init)
    return make_learner(cfg, init, mode=MODE_INORDER), p


@pytest.mark.parametrize("case", ["ragged_unsorted_dups", "two_field_k16", "long_unsorted_rows", "no_bias_no_linear",
                                  "k0", "empty_rows"])
def test_eterms_bit_exact(case, built_lib):
    k, k0, k1 = 8, 1, 1
    if case == "ragged_unsorted_dups":
        d = synth.ragged(20_000, 500, 11, seed=31)
    elif case == "two_field_k16":
        d, k = synth.two_field(200_000, 7000, 1000, seed=32), 16
    elif case == "long_unsorted_rows":   # > 64 entries: the selection path
        d, k = synth.ragged(300, 5000, 150, seed=33), 5
    elif case == "no_bias_no_linear":
        d, k0, k1 = synth.ragged(5000, 300, 6, seed=34), 0, 0
    elif case == "k0":
        d, k = synth.ragged(5000, 300, 6, seed=35), 0
    else:
        d = synth.ragged(3000, 100, 3, seed=36, empty_frac=0.8)
    l, p = _setup(d, k, k0, k1)
    got = l.mcmc_eterms(d)
    want = p.mcmc_eterms(d)
    assert np.array_equal(got, want)
    l.close()


def test_eterms_follow_set_params_and_refuse_fp32_state(built_lib):
    """the MCMC loop redraws the parameters on the host every iteration: set_params -> eterms"""
    d = synth.two_field(5000, 300, 200, seed=4)
    l, p = _setup(d, 4)
    r = np.random.default_rng(9)
    for _ in range(3):
        l.fm.w0 = float(r.standard_normal())
        l.fm.w = r.standard_normal(d.num_feature) * 0.2
        l.fm.v = r.standard_normal((4, d.num_feature)) * 0.2
        l.push_params()
        p.set_params(l.fm.w0, l.fm.w, l.fm.v)
        assert np.array_equal(l.mcmc_eterms(d), p.mcmc_eterms(d))
    l.set_mode(MODE_HOGWILD)
    with pytest.raises(FmError, match="fp64"):
        l.mcmc_eterms(d)
    l.close()


def test_eterms_c4_shape_full_size(built_lib):
    """BASELINE config C4 shape (MovieLens-10M: 10 000 054 cases, 71 567 users + 10 681 items, k = 16):
    the whole per-iteration re-prediction, bit-exact against the oracle on a 200k-case sample and
    checked on every case through a size-independent property (e-term == fp64 predict to rounding)."""
    d = synth.two_field(10_000_054, 71_567, 10_681, seed=5)
    l, p = _setup(d, 16)
    t0 = time.perf_counter()
    got = l.mcmc_eterms(d)
    wall = time.perf_counter() - t0
    sample = d.rows(4_000_000, 4_200_000)
    assert np.array_equal(got[4_000_000:4_200_000], p.mcmc_eterms(sample))
    pred = l.predict(d, transform=False)
    assert np.max(np.abs(got - pred)) < 1e-12
    t1 = time.perf_counter()
    p.mcmc_eterms(sample)
    cpu = (time.perf_counter() - t1) * d.num_cases / sample.num_cases
    print("\n[mcmc e-terms C4 shape] GPU incl. D2H of 80 MB: %.1f ms; oracle port on 1 core (scaled from 200k "
          "cases): %.0f ms" % (wall * 1e3, cpu * 1e3))
    l.close()
