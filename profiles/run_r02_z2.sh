#!/bin/bash
# Round 2, GPU call Z2: streaming traversal with bursts of 4 / 8 / 16 steps between the warp's meetings (RT3_BINNING=5) on C3, C2 and C5.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/z2_variants.jsonl
for n in 4 8 16; do
  lib=$PWD/profiles/librt3cuda_burst$n.so
  RT3_CORE_LIB=$lib RT3_BINNING=5 timeout 300 python profiles/variants.py burst-$n --c3 >> $OUT/z2_variants.jsonl 2>> $OUT/z2_variants.err
  RT3_CORE_LIB=$lib RT3_BINNING=5 timeout 300 python profiles/variants.py burst-$n --c3 --spp 16 >> $OUT/z2_variants.jsonl 2>> $OUT/z2_variants.err
  RT3_CORE_LIB=$lib RT3_BINNING=5 timeout 300 python profiles/variants.py burst-$n --c2bvh >> $OUT/z2_variants.jsonl 2>> $OUT/z2_variants.err
  RT3_CORE_LIB=$lib RT3_BINNING=5 timeout 300 python profiles/variants.py burst-$n --c5 >> $OUT/z2_variants.jsonl 2>> $OUT/z2_variants.err
done
