#!/bin/bash
# Round-2 GPU call M: A/B of ORDERED kernel variants on ONE box (alternative builds under libfm_b200/lib/alt).
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
for t in v8 h1 h2 h3; do
  echo "== $t"; FMB200_LIB=$PWD/libfm_b200/lib/alt/libfmb200_$t.so timeout 60 python scripts/prof_ordered.py 200000 0 0
done
echo "== head"; timeout 60 python scripts/prof_ordered.py 200000 0 0
done
echo "== head 1M"; timeout 60 python scripts/prof_ordered.py 1000209 0 0
echo "== v8 1M"; FMB200_LIB=$PWD/libfm_b200/lib/alt/libfmb200_v8.so timeout 60 python scripts/prof_ordered.py 1000209 0 0
echo "== h1 1M"; FMB200_LIB=$PWD/libfm_b200/lib/alt/libfmb200_h1.so timeout 60 python scripts/prof_ordered.py 1000209 0 0
} > gpurun_out/r2_ordered_ab.txt 2>&1
cut -c1-90 gpurun_out/r2_ordered_ab.txt
