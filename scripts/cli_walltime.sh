#!/bin/bash
# End-to-end wall time of the drop-in CLI vs the stock reference CLI on a C2-shaped text file
# (development aid; run on the GPU box from the repo root).
set -u
cd "$(dirname "$0")/.."
python - <<'PY'
import sys; sys.path.insert(0, ".")
from libfm_b200 import synth
d = synth.movielens_1m_shaped(seed=7, planted_k=4)
synth.to_libfm_text(d.rows(0, 900000), "/tmp/tr.libfm")
synth.to_libfm_text(d.rows(900000, d.num_cases), "/tmp/te.libfm")
PY
A="-task r -train /tmp/tr.libfm -test /tmp/te.libfm -method sgd -dim 1,1,8 -iter 20 -learn_rate 0.01 -seed 42"
run() { local name="$1"; shift; local t0=$(date +%s.%N); "$@" > /tmp/cli_out.txt 2>&1; local rc=$?; local t1=$(date +%s.%N);
        echo "$name: rc=$rc wall $(echo "$t1 - $t0" | bc) s :: $(grep -E '^Final|ERROR' /tmp/cli_out.txt | tr '\t' ' ')"; }
run "ours hogwild 20 iters" bin/libFM $A
run "ours inorder  2 iters" bin/libFM $A -mode inorder -iter 2
run "reference    20 iters" oracle/_ref/libFM $A
run "reference     2 iters" oracle/_ref/libFM $A -iter 2
