#!/bin/bash
# Round 2, GPU call T: threshold of the primitive-parallel tail (RT3_TAIL_RAYS = 0 (off), 4, 8 (default), 16, 32) on the whole C2 frame and
# on one rank's eighth of it (tile_rows 1, part 3 of 8), where the kernel's fixed tail shows.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/t_variants.jsonl
for round in 1 2; do
  for t in 0 4 8 16 32; do
    if [ $t = 8 ]; then lib=$PWD/raytracer-3_b200/csrc/librt3cuda.so; else lib=$PWD/profiles/librt3cuda_tail$t.so; fi
    RT3_CORE_LIB=$lib python profiles/variants.py tail-$t --eighth --reps 6 >> $OUT/t_variants.jsonl 2>> $OUT/t_variants.err
  done
done
