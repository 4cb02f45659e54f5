#!/bin/bash
# Round 2, GPU call C: binned hierarchy traversal (parity + C3 timing with and without it) and sweep-kernel occupancy variants.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/c_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/c_pytest.log
: > $OUT/c_variants.jsonl
python profiles/variants.py default --c3 >> $OUT/c_variants.jsonl 2>> $OUT/c_variants.err
RT3_NO_BINNING=1 python profiles/variants.py default-no-binning --c3 >> $OUT/c_variants.jsonl 2>> $OUT/c_variants.err
python profiles/variants.py default --c3 --spp 4 >> $OUT/c_variants.jsonl 2>> $OUT/c_variants.err
RT3_NO_BINNING=1 python profiles/variants.py default-no-binning --c3 --spp 4 >> $OUT/c_variants.jsonl 2>> $OUT/c_variants.err
python profiles/variants.py default >> $OUT/c_variants.jsonl 2>> $OUT/c_variants.err
for v in v1 v2 v3; do RT3_CORE_LIB=$PWD/profiles/librt3cuda_$v.so python profiles/variants.py $v >> $OUT/c_variants.jsonl 2>> $OUT/c_variants.err; done
timeout 300 python profiles/bvh_c3_probe.py > $OUT/c_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02c_bvh_c3 python profiles/bvh_c3_probe.py > $OUT/c_ncu_bvh.log 2>&1
ls -la $OUT > $OUT/c_listing.txt
