"""GPU: FMB200_MODE_ORDERED (libfm_b200/csrc/fm_ordered.cuh) -- the sequentially consistent epoch --
through the C ABI against the sequential oracle (oracle/fm_oracle.c, pinned to the reference).

The mode keeps the reference's read/write order on w0 / w / V and re-associates three sums (score
addends, bias chain by affine scan, FMA contraction), so the bar is NOT bit equality but a tolerance
far inside the north star's 1e-5 RMSE: parameters to 1e-9 relative, RMSE to 1e-9, asserted per
epoch -- including BASELINE config C2 at full size (the configuration the headline is quoted on).
"""
import time

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, make_learner
from libfm_b200 import MODE_INORDER, MODE_ORDERED, synth
from oracle import Port
from test_oracle import _ragged_short_rows

pytestmark = pytest.mark.gpu

RMSE_TOL = 1e-5   # BASELINE.json north_star tolerance
PARAM_RTOL = 1e-9  # what the mode actually delivers (rounding-level differences only)


def _cfg(n, k, task=0, lr=0.01, regs=(0, 0, 0), k0=1, k1=1, mn=1.0, mx=5.0):
    return dict(n=n, k=k, k0=k0, k1=k1, task=task, lr=lr, regs=np.array(regs, dtype=float),
                min_target=mn, max_target=mx)


def _port(cfg, init):
    p = Port(cfg["n"], cfg["k"], cfg["k0"], cfg["k1"])
    p.set_params(*init)
    p.reg0, p.regw, p.regv = [float(x) for x in cfg["regs"]]
    return p


def _rand_init(n, k, seed, stdev=0.1):
    r = np.random.default_rng(seed)
    return (float(r.standard_normal() * 0.1), r.standard_normal(n) * 0.1,
            r.standard_normal((k, n)) * stdev)


def _close(got, want, rtol=PARAM_RTOL):
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * 1e-3)


def _host_index(d):
    """restatement of the dependency index (same as tests/simt/ordered_host.cpp::build_links)"""
    NONE = 0xFFFFFFFF
    link = np.full(d.num_values, NONE, dtype=np.uint64)
    rowdep = np.full(d.num_cases, NONE, dtype=np.uint64)
    rows = np.repeat(np.arange(d.num_cases, dtype=np.int64), np.diff(d.row_ptr.astype(np.int64)))
    order = np.argsort(d.col, kind="stable")
    cs, es = d.col[order], order.astype(np.int64)
    same = np.zeros(len(order), dtype=bool)
    same[1:] = cs[1:] == cs[:-1]
    idx = np.nonzero(same)[0]
    link[es[idx]] = es[idx] - es[idx - 1]
    dist = rows[es[idx]] - rows[es[idx - 1]]
    np.minimum.at(rowdep, rows[es[idx]], dist.astype(np.uint64))
    return link.astype(np.uint32), rowdep.astype(np.uint32)


@pytest.mark.parametrize("case", ["two_field", "ragged_dups", "long_rows", "empty"])
def test_dependency_index_is_bit_exact(case, built_lib):
    if case == "two_field":
        d = synth.two_field(50_000, 600, 400, seed=3, zipf=1.0)
    elif case == "ragged_dups":
        d = _ragged_short_rows(30_000, 300, seed=6, dup_every=7)
    elif case == "long_rows":
        d = synth.ragged(3_000, 5_000, 60, seed=8)
    else:
        d = synth.ragged(200, 50, 3, seed=9, empty_frac=0.9)
    n = d.num_feature
    l = make_learner(_cfg(n, 4), _rand_init(n, 4, 1), mode=MODE_ORDERED)
    link, rowdep = l.ordered_index(d)
    want_link, want_rowdep = _host_index(d)
    assert np.array_equal(link, want_link)
    assert np.array_equal(rowdep, want_rowdep)
    l.close()


def _shape(case):
    task, k, k0, k1, regs, lr = 0, 8, 1, 1, (0.0, 0.0, 0.0), 0.02
    if case == "c2_shape":
        tr = synth.two_field(200_000, 6040, 3706, seed=3, planted_k=4)
    elif case == "zipf":
        tr = synth.two_field(40_000, 600, 400, seed=4, zipf=1.1)
    elif case == "ragged":
        tr = _ragged_short_rows(30_000, 500, seed=5)
    elif case == "dups":
        tr = _ragged_short_rows(30_000, 300, seed=6, dup_every=7)
        regs = (0.01, 0.02, 0.03)
    elif case == "classification":
        tr = synth.two_field(30_000, 800, 500, seed=7)
        tr.target[:] = np.where(tr.target > 3, 1.0, -1.0)
        task = 1
    elif case == "no_bias":
        tr = synth.two_field(30_000, 800, 500, seed=8)
        k0, k1 = 0, 0
    elif case == "two_field_real_values":
        # two entries per row but values != 1: the register-resident path WITHOUT the one-hot shortcut
        tr = synth.two_field(60_000, 1500, 900, seed=12, planted_k=4)
        tr.val[:] = (0.5 + np.random.default_rng(3).random(tr.val.shape[0])).astype(np.float32)
        regs = (0.0, 0.01, 0.02)
    elif case == "one_hot_regularised":
        # the one-hot two-field path with all three regularisers and a user id that is also an item id's twin
        tr = synth.two_field(60_000, 1200, 800, seed=13, planted_k=4)
        regs = (0.02, 0.01, 0.03)
    elif case == "clamp_heavy":
        # big steps push the scores across min/max_target all the time: the clamp guesses of the bias chain are
        # contradicted and the chain is re-walked (95 re-walks on the 700-row twin of this case in tests/simt)
        tr = synth.two_field(3_000, 70, 50, seed=14, planted_k=4)
        lr = 0.3
    elif case == "tiny":
        tr = synth.two_field(5, 3, 3, seed=1)
    elif case == "k3_reg":
        tr = _ragged_short_rows(20_000, 400, seed=9)
        k, regs = 3, (0.01, 0.02, 0.03)
    elif case == "k16_c4_shape":
        tr = synth.two_field(100_000, 7000, 1000, seed=10, planted_k=4)
        k = 16
    elif case == "k64_fields":
        tr = synth.multi_field(4_000, 39, 39 * 300, seed=11)
        tr.target[:] = np.where(tr.target > 0, 1.0, -1.0)
        task, k, lr = 1, 64, 0.01
    elif case == "k128_long":
        tr = synth.ragged(1_500, 2_000, 50, seed=12)
        k, lr = 128, 0.002
    else:
        raise ValueError(case)
    return tr, task, k, k0, k1, regs, lr


@pytest.mark.parametrize("case", ["tiny", "c2_shape", "zipf", "ragged", "dups", "classification", "no_bias",
                                  "k3_reg", "k16_c4_shape", "k64_fields", "k128_long", "two_field_real_values",
                                  "one_hot_regularised", "clamp_heavy"])
def test_ordered_matches_sequential_oracle(case, built_lib):
    tr, task, k, k0, k1, regs, lr = _shape(case)
    n = tr.num_feature
    init = _rand_init(n, k, 1)
    mn, mx = float(tr.target.min()), float(tr.target.max())
    cfg = _cfg(n, k, task=task, lr=lr, regs=regs, k0=k0, k1=k1, mn=mn, mx=mx)
    p = _port(cfg, init)
    l = make_learner(cfg, init, mode=MODE_ORDERED)
    for ep in range(2):
        sec = l.sgd_epoch(tr)
        p.sgd_epoch(tr, task, lr, mn, mx)
        got, want = l.evaluate(tr), p.metric(tr, task, mn, mx)
        assert abs(got - want) <= (1e-9 if task == 0 else 0.0), (case, ep, got, want)
    l.pull_params()
    if k0:
        _close(l.fm.w0, p.w0.value)
    _close(l.fm.w, p.w)
    _close(l.fm.v, p.v)
    cfgd = l.epoch_config()
    print("\n[ordered %s] %d rows k=%d: %.3f ms/epoch = %.2f M ex/s  %s" %
          (case, tr.num_cases, k, sec * 1e3, tr.num_cases / sec / 1e6, cfgd))
    l.close()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_ordered_matches_reference_golden(name, built_lib):
    """the vectors the reference itself produced (scripts/make_golden.py)"""
    z, tr, te = load_golden(name)
    l = make_learner(z, (float(z["w0_init"]), z["w_init"], z["v_init"]), mode=MODE_ORDERED)
    for e in range(int(z["epochs"])):
        l.sgd_epoch(tr)
        assert abs(l.evaluate(tr) - z["metric_train"][e]) <= 1e-9
        assert abs(l.evaluate(te) - z["metric_test"][e]) <= 1e-9
    l.pull_params()
    _close(l.fm.w0, float(z["w0"]))
    _close(l.fm.w, z["w"])
    _close(l.fm.v, z["v"])
    np.testing.assert_allclose(l.predict(te), z["pred_test"], rtol=1e-9, atol=1e-9)
    l.close()


def test_ordered_and_inorder_share_state(built_lib):
    """mode switches between the two fp64 modes keep the parameters bit for bit"""
    tr = synth.two_field(20_000, 600, 400, seed=2)
    n = tr.num_feature
    cfg = _cfg(n, 8)
    init = _rand_init(n, 8, 3)
    l = make_learner(cfg, init, mode=MODE_ORDERED)
    l.sgd_epoch(tr)
    l.pull_params()
    w0, w, v = l.fm.w0, l.fm.w.copy(), l.fm.v.copy()
    l.set_mode(MODE_INORDER)
    l.pull_params()
    assert l.fm.w0 == w0 and np.array_equal(l.fm.w, w) and np.array_equal(l.fm.v, v)
    p = _port(cfg, (w0, w, v))
    l.sgd_epoch(tr)  # exact kernel continues from the ordered state
    p.sgd_epoch(tr, 0, 0.01, 1.0, 5.0)
    l.pull_params()
    assert l.fm.w0 == p.w0.value and np.array_equal(l.fm.v, p.v)
    l.close()


@pytest.mark.parametrize("threads", [128, 256, 1024])
def test_ordered_thread_counts(threads, built_lib):
    tr = synth.two_field(60_000, 2000, 1500, seed=12, planted_k=4)
    n = tr.num_feature
    cfg = _cfg(n, 8, lr=0.02)
    init = _rand_init(n, 8, 5)
    p = _port(cfg, init)
    l = make_learner(cfg, init, mode=MODE_ORDERED)
    l.set_tuning(threads=threads)
    l.sgd_epoch(tr)
    p.sgd_epoch(tr, 0, 0.02, 1.0, 5.0)
    l.pull_params()
    _close(l.fm.w0, p.w0.value)
    _close(l.fm.v, p.v)
    l.close()


def test_ordered_c2_full_size_trajectory(built_lib):
    """BASELINE config C2 at full size (1 000 209 rows, 6040 x 3706, k = 8), planted signal, 5 epochs:
    |RMSE_gpu - RMSE_oracle| <= 1e-5 every epoch (asserted 1e-8), parameters to 1e-9, and the epoch is
    timed next to the oracle's single core (report only; bench.py carries the driver-run number)."""
    tr, te = synth.movielens_1m_planted(100_000, seed=7)
    n = tr.num_feature
    init = _rand_init(n, 8, 42)
    init = (0.0, np.zeros(n), init[2])
    cfg = _cfg(n, 8, lr=0.01, mn=tr.min_target, mx=tr.max_target)
    p = _port(cfg, init)
    l = make_learner(cfg, init, mode=MODE_ORDERED)
    gaps, secs, cpu = [], [], []
    for ep in range(5):
        secs.append(l.sgd_epoch(tr))
        t0 = time.perf_counter()
        p.sgd_epoch(tr, 0, 0.01, tr.min_target, tr.max_target)
        cpu.append(time.perf_counter() - t0)
        g_tr, g_te = l.evaluate(tr), l.evaluate(te)
        o_tr = p.metric(tr, 0, tr.min_target, tr.max_target)
        o_te = p.metric(te, 0, tr.min_target, tr.max_target)
        gaps.append(max(abs(g_tr - o_tr), abs(g_te - o_te)))
        assert gaps[-1] <= 1e-8 <= RMSE_TOL, (ep, g_tr, o_tr, g_te, o_te)
    l.pull_params()
    _close(l.fm.w0, p.w0.value)
    _close(l.fm.w, p.w)
    _close(l.fm.v, p.v)
    print("\n[ordered C2 full] %.2f ms/epoch = %.1f M ex/s (oracle port, 1 core here: %.1f M ex/s); "
          "max RMSE gap over 5 epochs %.2e; %s" %
          (min(secs) * 1e3, tr.num_cases / min(secs) / 1e6, tr.num_cases / min(cpu) / 1e6, max(gaps),
           l.epoch_config()))
    l.close()


def test_inorder_c2_full_bit_exact(built_lib):
    """in-order parity at C2 size (one epoch, 1.4 s): bit-exact parameters."""
    tr = synth.movielens_1m_shaped(seed=7, planted_k=4)
    n = tr.num_feature
    init = _rand_init(n, 8, 42)
    cfg = _cfg(n, 8, lr=0.01, mn=tr.min_target, mx=tr.max_target)
    p = _port(cfg, init)
    l = make_learner(cfg, init, mode=MODE_INORDER)
    l.sgd_epoch(tr)
    p.sgd_epoch(tr, 0, 0.01, tr.min_target, tr.max_target)
    l.pull_params()
    assert l.fm.w0 == p.w0.value and np.array_equal(l.fm.w, p.w) and np.array_equal(l.fm.v, p.v)
    assert abs(l.evaluate(tr) - p.metric(tr, 0, tr.min_target, tr.max_target)) <= 1e-12
    l.close()
