#!/bin/bash
# Round 2, GPU call Z: streaming traversal (RT3_BINNING=5: queue of fresh rays per warp, idle lanes refill, long walks carried into the next round)
# against the three-class per-warp sort (3, the default) on C3, and against the plain traversal (default for sphere scenes) on C2 and C5.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/z_variants.jsonl
RT3_BINNING=5 timeout 600 python -m pytest tests -m gpu -x -q -k "bvh or hierarchy or full_size" > $OUT/z_pytest_b5.log 2>&1; echo "pytest rc=$?" >> $OUT/z_pytest_b5.log
for b in 3 5; do
  RT3_BINNING=$b timeout 300 python profiles/variants.py binning-$b --c3 --spp 16 >> $OUT/z_variants.jsonl 2>> $OUT/z_variants.err
  RT3_BINNING=$b timeout 300 python profiles/variants.py binning-$b --c3 >> $OUT/z_variants.jsonl 2>> $OUT/z_variants.err
done
for b in 0 5; do
  RT3_BINNING=$b timeout 300 python profiles/variants.py binning-$b --c2bvh >> $OUT/z_variants.jsonl 2>> $OUT/z_variants.err
  RT3_BINNING=$b timeout 300 python profiles/variants.py binning-$b --c5 >> $OUT/z_variants.jsonl 2>> $OUT/z_variants.err
done
RT3_BINNING=5 timeout 300 python profiles/soak.py 300 7021 > $OUT/z_soak_b5.log 2>&1; echo "rc=$?" >> $OUT/z_soak_b5.log
