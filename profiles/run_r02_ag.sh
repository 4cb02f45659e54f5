#!/bin/bash
# Round 2, GPU call AG: primary rays through the candidates of a beam's walk of the hierarchy (ACCEL + BEAM kernels). Parity first (the
# hierarchy tests, the full-shape bands against the oracle, soak rows at many samples per pixel), then C5 / C3 / C2 through the hierarchy with
# and without (RT3_BEAM_BVH=0).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 700 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_full_size.py tests/test_gpu_pathtrace.py -x -q -m gpu > $OUT/ag_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ag_pytest.log
timeout 300 python profiles/soak.py 400 9101 > $OUT/ag_soak.log 2>&1; echo "rc=$?" >> $OUT/ag_soak.log
: > $OUT/ag_variants.jsonl
for b in 0 1; do
  RT3_BEAM_BVH=$b timeout 200 python profiles/variants.py beam-bvh-$b --c5 >> $OUT/ag_variants.jsonl 2>> $OUT/ag_variants.err
  RT3_BEAM_BVH=$b timeout 200 python profiles/variants.py beam-bvh-$b --c3 >> $OUT/ag_variants.jsonl 2>> $OUT/ag_variants.err
  RT3_BEAM_BVH=$b timeout 200 python profiles/variants.py beam-bvh-$b --c3 --spp 32 >> $OUT/ag_variants.jsonl 2>> $OUT/ag_variants.err
  RT3_BEAM_BVH=$b timeout 200 python profiles/variants.py beam-bvh-$b --c2bvh >> $OUT/ag_variants.jsonl 2>> $OUT/ag_variants.err
done
tail -5 $OUT/ag_pytest.log; tail -3 $OUT/ag_soak.log; cat $OUT/ag_variants.jsonl | cut -c1-400
