#!/bin/bash
# Round 2, GPU call Z3: hierarchy traversal with (pf1 / pf2) a prefetch of the stacked subtree's record into L1 / L2 and (eager) the triangle's
# corners requested together with its plane, against the committed build, on C3 (256 and 16 spp), C2 and C5.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/z3_variants.jsonl
for v in base pf1 pf2 eager eager_pf1 base; do
  if [ $v = base ]; then lib=$PWD/raytracer-3_b200/csrc/librt3cuda.so; else lib=$PWD/profiles/librt3cuda_$v.so; fi
  RT3_CORE_LIB=$lib timeout 300 python profiles/variants.py $v --c3 >> $OUT/z3_variants.jsonl 2>> $OUT/z3_variants.err
  RT3_CORE_LIB=$lib timeout 300 python profiles/variants.py $v --c3 --spp 16 >> $OUT/z3_variants.jsonl 2>> $OUT/z3_variants.err
  RT3_CORE_LIB=$lib timeout 300 python profiles/variants.py $v --c2bvh >> $OUT/z3_variants.jsonl 2>> $OUT/z3_variants.err
  RT3_CORE_LIB=$lib timeout 300 python profiles/variants.py $v --c5 >> $OUT/z3_variants.jsonl 2>> $OUT/z3_variants.err
done
