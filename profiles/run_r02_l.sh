#!/bin/bash
# Round 2, GPU call L: one 64-byte record per face (normal + plane offset + three vertices) instead of four arrays: parity and C3 timing
# against the four-array build (profiles/librt3cuda_soa.so = the previous commit).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/l_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/l_pytest.log
: > $OUT/l_variants.jsonl
for spp in 256 4; do
  python profiles/variants.py face-records --c3 --spp $spp >> $OUT/l_variants.jsonl 2>> $OUT/l_variants.err
  RT3_CORE_LIB=$PWD/profiles/librt3cuda_soa.so python profiles/variants.py four-arrays --c3 --spp $spp >> $OUT/l_variants.jsonl 2>> $OUT/l_variants.err
done
python profiles/variants.py face-records >> $OUT/l_variants.jsonl 2>> $OUT/l_variants.err
timeout 600 python profiles/configs.py c3 > $OUT/l_configs.jsonl 2> $OUT/l_configs.err
timeout 300 python profiles/soak.py 1500 7008 > $OUT/l_soak.log 2>&1
timeout 300 python profiles/sweep_rate.py 4096 65536 > $OUT/l_sweep_rate.jsonl 2>&1
