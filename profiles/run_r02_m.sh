#!/bin/bash
# Round 2, GPU call M: the committed tree the way the driver runs it at round end: smoke, the GPU suite, the reference arm, the bench line.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/m_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/m_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/m_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/m_pytest.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/m_bench_reference_arm.json 2> $OUT/m_bench.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/m_bench.json 2>> $OUT/m_bench.err; echo "bench rc=$?" >> $OUT/m_bench.err
timeout 300 python profiles/soak.py 1000 7009 > $OUT/m_soak.log 2>&1
