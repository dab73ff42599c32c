#!/bin/bash
# Round-2 GPU call O: ORDERED v12 (Kogge-Stone threading of the segment totals, one-FMA update without regularisation)
# against v11 on one box.
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== v11"; FMB200_LIB=$PWD/libfm_b200/lib/alt/libfmb200_v11.so timeout 60 python scripts/prof_ordered.py 200000 0 0
echo "== v12"; timeout 60 python scripts/prof_ordered.py 200000 0 0
done
echo "== v12 phases"; timeout 60 python scripts/prof_ordered.py 200000 0 132
echo "== v12 1M"; timeout 60 python scripts/prof_ordered.py 1000209 0 0
} > gpurun_out/r2_ordered_v12.txt 2>&1
cut -c1-420 gpurun_out/r2_ordered_v12.txt
timeout 300 python -m pytest tests/test_ordered_gpu.py -m gpu -q -x 2>&1 | tail -3
