#!/bin/bash
# Round 2, GPU call AB: primary rays through a per-chunk candidate list (BEAM kernel) against the plain sweep (RT3_BEAM=0) on C2:
# kernel times on the whole frame and on one rank's eighth of it, parity (GPU suite, soak), ncu capture of the new kernel.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/ab_variants.jsonl
for b in 0 1 0 1; do
  RT3_BEAM=$b timeout 300 python profiles/variants.py beam-$b --eighth --reps 4 >> $OUT/ab_variants.jsonl 2>> $OUT/ab_variants.err
done
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/ab_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ab_pytest.log
timeout 300 python profiles/soak.py 1500 7041 > $OUT/ab_soak.log 2>&1; echo "rc=$?" >> $OUT/ab_soak.log
bash profiles/run_r02_ac.sh
