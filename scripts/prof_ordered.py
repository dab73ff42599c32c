"""ncu target: a few ORDERED epochs on a C2-shaped set (development aid).
    ncu --set full --import-source on -k regex:ordered -s 1 -c 1 -o gpurun_out/r2_ordered python scripts/prof_ordered.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from libfm_b200 import MODE_ORDERED, FmLearnSgdElement, FmModel, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # >= 100: timing experiments (wrong results)
d = synth.two_field(rows, 6040, 3706, seed=3, planted_k=4)
fm = FmModel(d.num_feature, 8)
fm.init_stdev = 0.1
fm.init_numpy(42)
l = FmLearnSgdElement(fm, mode=MODE_ORDERED)
l.task, l.learn_rate = 0, 0.01
l.min_target, l.max_target = d.min_target, d.max_target
l.push_hparams()
if threads or variant:
    l.set_tuning(threads=threads, variant=variant)
for _ in range(3):
    t = l.sgd_epoch(d)
print("ordered rows=%d threads=%d variant=%d: epoch %.3f ms = %.1f M ex/s %s" % (
    rows, threads, variant, t * 1e3, rows / t / 1e6, l.epoch_config()))
l.close()
