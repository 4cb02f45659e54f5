#!/bin/bash
# Round 2, GPU call D: the hierarchy kernels without the in-loop overflow atomic (call C showed +40 % with it), binned vs unbinned.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/d_variants.jsonl
for spp in 256 32 4; do
  python profiles/variants.py binned --c3 --spp $spp >> $OUT/d_variants.jsonl 2>> $OUT/d_variants.err
  RT3_NO_BINNING=1 python profiles/variants.py unbinned --c3 --spp $spp >> $OUT/d_variants.jsonl 2>> $OUT/d_variants.err
done
timeout 600 python profiles/configs.py c2 c5 > $OUT/d_configs.jsonl 2> $OUT/d_configs.err
timeout 900 python -m pytest tests -m gpu -x -q -k "bvh or hierarchy or full_size or book or stress" > $OUT/d_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/d_pytest.log
