#!/bin/bash
# Round 2, GPU call R: the primitive-parallel tail of the path tracer (sweep_slots_by_primitive): parity, soak, bench.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r_pytest.log
timeout 300 python profiles/soak.py 2000 7012 > $OUT/r_soak.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r_bench.json 2> $OUT/r_bench.err; echo "bench rc=$?" >> $OUT/r_bench.err
timeout 300 python profiles/configs.py c1 > $OUT/r_configs.jsonl 2>&1
