#!/bin/bash
# Round 2, GPU call AD: BEAM kernel (primary rays traced against a per-chunk candidate list and shaded at regeneration) with 1 / 2 / 3 / 4
# regeneration batches per round against the plain sweep (RT3_BEAM=0), C2 whole frame and one rank's eighth.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/ad_variants.jsonl
RT3_BEAM=0 timeout 300 python profiles/variants.py beam-off --eighth --reps 4 >> $OUT/ad_variants.jsonl 2>> $OUT/ad_variants.err
for n in 1 2 3 4; do
  if [ $n = 2 ]; then lib=$PWD/raytracer-3_b200/csrc/librt3cuda.so; else lib=$PWD/profiles/librt3cuda_batches$n.so; fi
  RT3_CORE_LIB=$lib timeout 300 python profiles/variants.py beam-batches-$n --eighth --reps 4 >> $OUT/ad_variants.jsonl 2>> $OUT/ad_variants.err
done
