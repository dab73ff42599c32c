"""Where does the C2 epoch's time go?  (development aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libfm_b200 import FmLearnSgdElement, FmModel, MODE_HOGWILD, synth

def run(name, d, k, k0=True, k1=True, tune=(0,0,0,-1,0), epochs=6):
    fm = FmModel(d.num_feature, k, k0, k1); fm.init_stdev = 0.1; fm.init_numpy(42)
    l = FmLearnSgdElement(fm, mode=MODE_HOGWILD)
    l.task, l.learn_rate = 0, 0.01
    l.min_target, l.max_target = d.min_target, d.max_target
    l.push_hparams(); l.set_tuning(*tune)
    ts = [l.sgd_epoch(d) for _ in range(epochs)]
    print("%-28s dbg=%s best %.1f us cfg=%s" % (name, os.environ.get("FMB200_DEBUG", "0"), min(ts[1:]) * 1e6, l.epoch_config()), flush=True)
    l.close()

d = synth.movielens_1m_shaped(seed=7)
for dbg in ("0", "1", "2", "3"):
    os.environ["FMB200_DEBUG"] = dbg
    run("C2 full", d, 8)
os.environ["FMB200_DEBUG"] = "0"
run("C2 k1=0 (no w)", d, 8, k1=False)
run("C2 k0=0 (no bias)", d, 8, k0=False)
run("C2 k0=k1=0", d, 8, k0=False, k1=False)
os.environ["FMB200_DEBUG"] = "1"
run("C2 k0=k1=0 noVred", d, 8, k0=False, k1=False)
os.environ["FMB200_DEBUG"] = "0"
# spread the same rows over 16x more features: fewer same-line collisions
import numpy as np
d2 = synth.two_field(1_000_209, 6040 * 16, 3706 * 16, seed=7)
run("C2 16x features", d2, 8)
os.environ["FMB200_DEBUG"] = "3"
run("C2 16x features noRED", d2, 8)
