"""A few hogwild epochs of one workload, for ncu (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libfm_b200 import FmLearnSgdElement, FmModel, MODE_HOGWILD, synth

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if which == "c2":
    d, k, task = synth.movielens_1m_shaped(seed=7), 8, 0
elif which == "c3":
    d, k, task = synth.multi_field(1_000_000, 39, 1_000_000, 11), 64, 1
    d.binarize_targets()
fm = FmModel(d.num_feature, k); fm.init_stdev = 0.1; fm.init_numpy(42)
l = FmLearnSgdElement(fm, mode=MODE_HOGWILD)
l.task, l.learn_rate = task, 0.01
l.min_target, l.max_target = d.min_target, d.max_target
l.push_hparams()
for _ in range(n_epochs):
    t = l.sgd_epoch(d)
print(which, "last epoch us", t * 1e6, l.epoch_config())
