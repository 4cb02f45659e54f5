#!/bin/bash
# Round 2, GPU call E: no sort / CTA-wide sort / per-warp sort of the rays in front of the hierarchy traversal, C3 at 256, 16 and 4 spp.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/e_variants.jsonl
for spp in 256 16 4; do
  for b in 0 1 2; do
    RT3_BINNING=$b python profiles/variants.py binning-$b --c3 --spp $spp >> $OUT/e_variants.jsonl 2>> $OUT/e_variants.err
  done
done
timeout 600 python -m pytest tests -m gpu -x -q -k "bvh or hierarchy" > $OUT/e_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/e_pytest.log
RT3_BINNING=2 timeout 600 python -m pytest tests -m gpu -x -q -k "bvh or hierarchy or full_size" > $OUT/e_pytest_b2.log 2>&1; echo "pytest rc=$?" >> $OUT/e_pytest_b2.log
