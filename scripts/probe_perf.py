"""Quick on-GPU timing probe (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libfm_b200 import FmLearnSgdElement, FmModel, MODE_HOGWILD, MODE_INORDER, synth

def run(name, d, k, task=0, tunings=((0,0,0,0,0),), epochs=5, inorder=False):
    n = d.num_feature
    fm = FmModel(n, k); fm.init_stdev = 0.1; fm.init_numpy(42)
    l = FmLearnSgdElement(fm, mode=MODE_HOGWILD)
    l.task, l.learn_rate = task, 0.01
    l.min_target, l.max_target = d.min_target, d.max_target
    l.push_hparams()
    l.upload(d, 0)
    for t in tunings:
        l.set_tuning(*t)
        ts = [l.sgd_epoch(d) for _ in range(epochs)]
        best = min(ts[1:])
        bytes_ex = 2 * k * (d.num_values / d.num_cases) * 4
        print("%s tune=%s cfg=%s best %.1f us  %.2f Gex/s  alg %.0f GB/s  metric %.4f" % (
            name, t, l.epoch_config(), best * 1e6, d.num_cases / best / 1e9,
            d.num_cases * bytes_ex / best / 1e9, l.evaluate(d)), flush=True)
    t0 = time.time(); l.evaluate(d); print("  evaluate wall %.1f ms" % ((time.time() - t0) * 1e3))
    if inorder:
        l.set_mode(MODE_INORDER)
        t = l.sgd_epoch(d)
        print("  inorder epoch %.1f ms (%.2f Mex/s)" % (t * 1e3, d.num_cases / t / 1e6))
    l.close()

if __name__ == "__main__":
    d = synth.movielens_1m_shaped(seed=7)
    tun = [(0,0,0,0,0), (0,0,0,0,3), (0,0,128,0,3), (4,0,0,0,3), (0,0,0,0,1)]
    run("C2", d, 8, tunings=tun, inorder=True)
    dz = synth.movielens_1m_shaped(seed=7, zipf=1.0)
    run("C2zipf", dz, 8, tunings=[(0,0,0,0,0),(0,0,0,0,3)])
    d3 = synth.multi_field(1_000_000, 39, 1_000_000, 11); d3.binarize_targets()
    run("C3-1M", d3, 64, task=1)
    run("k128", d3, 128, task=1)
    d4 = synth.two_field(2_000_000, 71567, 10681, 5)
    run("C4shape-k16", d4, 16)
