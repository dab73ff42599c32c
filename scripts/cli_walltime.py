"""End-to-end wall time of the drop-in CLI vs the stock reference CLI on a C2-shaped text
file (development aid; run on the GPU box from the repo root)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libfm_b200 import synth
d = synth.movielens_1m_shaped(seed=7, planted_k=4)
synth.to_libfm_text(d.rows(0, 900000), "/tmp/tr.libfm")
synth.to_libfm_text(d.rows(900000, d.num_cases), "/tmp/te.libfm")
base = ["-task", "r", "-train", "/tmp/tr.libfm", "-test", "/tmp/te.libfm", "-method", "sgd", "-dim", "1,1,8",
        "-learn_rate", "0.01", "-seed", "42"]
for name, exe, extra in [("ours hogwild 20 iters", "bin/libFM", ["-iter", "20", "-verbosity", "1"]),
                         ("ours inorder  2 iters", "bin/libFM", ["-iter", "2", "-mode", "inorder"]),
                         ("reference    20 iters", "oracle/_ref/libFM", ["-iter", "20"]),
                         ("reference     2 iters", "oracle/_ref/libFM", ["-iter", "2"])]:
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, exe)] + base + extra, capture_output=True, text=True)
    dt = time.time() - t0
    fin = [l for l in (r.stdout + r.stderr).splitlines() if l.startswith(("Final", "time:")) or "ERROR" in l]
    print("%s: rc=%d wall %.2f s :: %s" % (name, r.returncode, dt, " ".join(fin).replace("\t", " ")), flush=True)
