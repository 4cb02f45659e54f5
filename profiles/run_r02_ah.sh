#!/bin/bash
# Round 2, GPU call AH: candidate tests of the hierarchy's beams four at a time (loads together) against one at a time, candidate lists of 192
# against 128, on C5 and on C2 through the hierarchy; hierarchy parity tests of the new default (beams for the unsorted hierarchy kernel only);
# ncu capture of the C5 render.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_bvh.py -x -q -m gpu > $OUT/ah_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ah_pytest.log
: > $OUT/ah_variants.jsonl
V="python profiles/variants.py"
timeout 200 $V group4-cand192 --c5 >> $OUT/ah_variants.jsonl 2>> $OUT/ah_variants.err
timeout 200 $V group4-cand192 --c2bvh >> $OUT/ah_variants.jsonl 2>> $OUT/ah_variants.err
RT3_CORE_LIB=$PWD/profiles/librt3cuda_group0.so timeout 200 $V group1-cand192 --c5 >> $OUT/ah_variants.jsonl 2>> $OUT/ah_variants.err
RT3_CORE_LIB=$PWD/profiles/librt3cuda_group0.so timeout 200 $V group1-cand192 --c2bvh >> $OUT/ah_variants.jsonl 2>> $OUT/ah_variants.err
RT3_CORE_LIB=$PWD/profiles/librt3cuda_cand128.so timeout 200 $V group4-cand128 --c5 >> $OUT/ah_variants.jsonl 2>> $OUT/ah_variants.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pathtrace_kernel -c 1 -f -o $OUT/r02ah_bvh_beam_c5 $V ncu --c5 > $OUT/ah_ncu.log 2>&1
echo "ncu rc=$?" >> $OUT/ah_ncu.log
tail -3 $OUT/ah_pytest.log; cut -c1-330 $OUT/ah_variants.jsonl; tail -2 $OUT/ah_ncu.log
