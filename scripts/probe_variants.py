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
for dbg in ("0", "4", "8", "12"):
    os.environ["FMB200_DEBUG"] = dbg
    run("C2 full (4=no w0 RED, 8=no w0 load)", d, 8)
os.environ["FMB200_DEBUG"] = "0"
run("C2 k0=0", d, 8, k0=False)
