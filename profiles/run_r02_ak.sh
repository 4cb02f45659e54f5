#!/bin/bash
# Round 2, GPU call AK: a further regeneration batch of a round only for at least 1 (as before) / 4 / 8 / 16 free slots (RT3_BEAM_MIN_BATCH), C2 whole
# frame and one rank's eighth.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/ak_variants.jsonl
V="python profiles/variants.py"
timeout 200 $V minbatch-1 --eighth --reps 4 >> $OUT/ak_variants.jsonl 2>> $OUT/ak_variants.err
for m in 4 8 16; do
  RT3_CORE_LIB=$PWD/profiles/librt3cuda_minbatch$m.so timeout 200 $V minbatch-$m --eighth --reps 4 >> $OUT/ak_variants.jsonl 2>> $OUT/ak_variants.err
done
timeout 200 $V minbatch-1-again --eighth --reps 4 >> $OUT/ak_variants.jsonl 2>> $OUT/ak_variants.err
cut -c1-420 $OUT/ak_variants.jsonl
