#!/bin/bash
# Round-2 GPU call N: A/B of helper-warp placement for the ORDERED kernel on ONE box.
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
for t in v8 g0 g3 g4 g2; do
  echo "== $t"; FMB200_LIB=$PWD/libfm_b200/lib/alt/libfmb200_$t.so timeout 60 python scripts/prof_ordered.py 200000 0 0
done
echo "== head (64 parked + 64 helpers)"; timeout 60 python scripts/prof_ordered.py 200000 0 0
done
echo "== head 1M"; timeout 60 python scripts/prof_ordered.py 1000209 0 0
echo "== head phases"; timeout 60 python scripts/prof_ordered.py 200000 0 132
echo "== g2 1M"; FMB200_LIB=$PWD/libfm_b200/lib/alt/libfmb200_g2.so timeout 60 python scripts/prof_ordered.py 1000209 0 0
} > gpurun_out/r2_ordered_ab2.txt 2>&1
cut -c1-95 gpurun_out/r2_ordered_ab2.txt
timeout 300 python -m pytest tests/test_ordered_gpu.py -m gpu -q -x 2>&1 | tail -3
