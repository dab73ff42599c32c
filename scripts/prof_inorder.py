"""One in-order epoch on 20k C2-shaped rows, for ncu (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libfm_b200 import FmLearnSgdElement, FmModel, MODE_INORDER, synth
d = synth.movielens_1m_shaped(seed=7, n_rows=20000)
fm = FmModel(d.num_feature, 8); fm.init_stdev = 0.1; fm.init_numpy(42)
l = FmLearnSgdElement(fm, mode=MODE_INORDER)
l.task, l.learn_rate = 0, 0.01
l.min_target, l.max_target = d.min_target, d.max_target
l.push_hparams()
for _ in range(2):
    t = l.sgd_epoch(d)
print("inorder us/row", t / d.num_cases * 1e6)
