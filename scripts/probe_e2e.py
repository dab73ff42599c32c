"""Break down the host-buffer (e2e) step: upload / epoch / get_params (development aid)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libfm_b200 import FmLearnSgdElement, FmModel, MODE_HOGWILD, synth
d = synth.movielens_1m_shaped(seed=7)
fm = FmModel(d.num_feature, 8); fm.init_stdev = 0.1; fm.init_numpy(42)
l = FmLearnSgdElement(fm, mode=MODE_HOGWILD)
l.task, l.learn_rate = 0, 0.01
l.min_target, l.max_target = d.min_target, d.max_target
l.push_hparams()
lib, ctx = l.lib, l._ctx
P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
def stage(pin):
    out = []
    for a in (d.row_ptr, d.col, d.val, d.target):
        if pin:
            from libfm_b200.model import pinned_copy
            out.append((pinned_copy(a), None))
        else:
            out.append((a.copy(), None))
    return out
for pin in (False, True):
    bufs = stage(pin)
    rp, col, val, tgt = [b[0] for b in bufs]
    w0 = C.c_double(); w = np.empty(d.num_feature); v = np.empty((8, d.num_feature))
    def up(): assert lib.fmb200_upload_data(ctx, 0, d.num_cases, int(rp[-1]), P(rp, C.c_uint64), P(col, C.c_uint32), P(val, C.c_float), P(tgt, C.c_float)) == 0
    def ep(): assert lib.fmb200_sgd_epoch(ctx, 0, None) == 0
    def gp(): assert lib.fmb200_get_params(ctx, C.byref(w0), P(w, C.c_double), P(v, C.c_double)) == 0
    for name, fn in (("upload", up), ("epoch", ep), ("get_params", gp)):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(10): fn()
        print("pinned=%s %-10s %.3f ms" % (pin, name, (time.perf_counter() - t0) / 10 * 1e3), flush=True)
    t0 = time.perf_counter()
    for _ in range(10): up(); ep(); gp()
    dt = (time.perf_counter() - t0) / 10
    print("pinned=%s full step  %.3f ms -> %.2f G ex/s" % (pin, dt * 1e3, d.num_cases / dt / 1e9), flush=True)
