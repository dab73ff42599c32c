"""CPU: pin the plain-C restatement (oracle/fm_oracle.c) against (a) the golden
vectors the REFERENCE produced (tests/golden, scripts/make_golden.py) and
(b) the reference itself when its shim is available (oracle/_ref)."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden
from libfm_b200 import synth
from oracle import Port, Ref, have_ref


def _port_from_golden(z):
    p = Port(int(z["n"]), int(z["k"]), int(z["k0"]), int(z["k1"]))
    p.set_params(float(z["w0_init"]), z["w_init"], z["v_init"])
    p.reg0, p.regw, p.regv = [float(x) for x in z["regs"]]
    return p


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_init_matches_reference_draw_order(name):
    z, tr, te = load_golden(name)
    p = Port(int(z["n"]), int(z["k"]))
    p.init(int(z["seed"]), 0.0, float(z["init_stdev"]))
    assert p.w0.value == 0.0 and not p.w.any()
    assert np.array_equal(p.v, z["v_init"])  # bit-exact: same rand() stream, same order


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_epochs_bit_exact_vs_golden(name):
    z, tr, te = load_golden(name)
    p = _port_from_golden(z)
    task, lr = int(z["task"]), float(z["lr"])
    mn, mx = float(z["min_target"]), float(z["max_target"])
    for e in range(int(z["epochs"])):
        p.sgd_epoch(tr, task, lr, mn, mx)
        assert p.metric(tr, task, mn, mx) == z["metric_train"][e]
        assert p.metric(te, task, mn, mx) == z["metric_test"][e]
    assert p.w0.value == float(z["w0"])
    assert np.array_equal(p.w, z["w"])
    assert np.array_equal(p.v, z["v"])
    assert np.array_equal(p.predict(te, task, mn, mx, True), z["pred_test"])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_port_vs_live_reference_c1_shape():
    tr = synth.plumbing_10k()
    te = synth.plumbing_10k(seed=99, n_rows=2000)
    n, k = max(tr.num_feature, te.num_feature), 8
    ref = Ref(n, k, seed=42, init_stdev=0.1)
    port = Port(n, k)
    port.init(42, 0.0, 0.1)
    ref.learn(tr, te, 0, 0.01, 2, tr.min_target, tr.max_target)
    for _ in range(2):
        port.sgd_epoch(tr, 0, 0.01, tr.min_target, tr.max_target)
    w0, w, v = ref.get_params()
    assert w0 == port.w0.value and np.array_equal(w, port.w) and np.array_equal(v, port.v)
    assert ref.evaluate(te, 0, tr.min_target, tr.max_target) == port.metric(te, 0, tr.min_target, tr.max_target)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_port_rng_matches_reference():
    L = Ref.lib()
    p = Port(1, 1)
    L.ref_srand(C.c_long(123))
    a = [L.ref_ran_gaussian() for _ in range(1000)]
    p.lib.fmo_srand(C.c_long(123))
    b = [p.lib.fmo_ran_gaussian() for _ in range(1000)]
    assert a == b


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_port_predict_row_edge_cases():
    n, k = 20, 3
    ref = Ref(n, k, seed=5)
    w0, w, v = ref.get_params()
    w = np.linspace(-1, 1, n)
    ref.set_params(0.25, w, v)
    port = Port(n, k)
    port.set_params(0.25, w, v)
    for col, val in [([], []), ([3], [2.5]), ([3, 3], [1.0, -2.0]), ([0, 19, 7], [0.5, 1.5, -1.0])]:
        pr, sr, ssr = ref.predict_row(col, val)
        pp, sp, ssp = port.predict_row(col, val)
        assert pr == pp and np.array_equal(sr, sp) and np.array_equal(ssr, ssp)
