#!/bin/bash
# Round 2, GPU call AI: candidate tests that leave at a negative discriminant (exact_sphere_path_coherent) -- in the hierarchy's beams (main
# build) and, as a variant, in the sweep's BEAM kernel; parity of the new default first.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_pathtrace.py tests/test_gpu_full_size.py -x -q -m gpu > $OUT/ai_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ai_pytest.log
: > $OUT/ai_variants.jsonl
V="python profiles/variants.py"
timeout 200 $V early-out --c5 >> $OUT/ai_variants.jsonl 2>> $OUT/ai_variants.err
timeout 200 $V early-out --c2bvh >> $OUT/ai_variants.jsonl 2>> $OUT/ai_variants.err
timeout 200 $V sweep-plain --reps 4 >> $OUT/ai_variants.jsonl 2>> $OUT/ai_variants.err
RT3_CORE_LIB=$PWD/profiles/librt3cuda_sweep_early.so timeout 200 $V sweep-early-out --reps 4 >> $OUT/ai_variants.jsonl 2>> $OUT/ai_variants.err
timeout 200 $V sweep-plain-again --reps 4 >> $OUT/ai_variants.jsonl 2>> $OUT/ai_variants.err
RT3_CORE_LIB=$PWD/profiles/librt3cuda_sweep_early.so timeout 200 python profiles/soak.py 300 4242 > $OUT/ai_soak_sweep_early.log 2>&1
tail -3 $OUT/ai_pytest.log; cut -c1-330 $OUT/ai_variants.jsonl; tail -1 $OUT/ai_soak_sweep_early.log
