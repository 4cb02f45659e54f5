#!/bin/bash
# Round 2, GPU call I: leaner per-word bookkeeping in the level-1 loop (full words without the partial-word test, add-with-carry for the
# "word not empty" summary): parity, sweep rates, bench, and the ncu capture of the intersection loop for the FMA-pipe share.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/i_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/i_pytest.log
timeout 300 python profiles/sweep_rate.py 484 4096 65536 1048576 > $OUT/i_sweep_rate.jsonl 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/i_bench.json 2> $OUT/i_bench.err; echo "bench rc=$?" >> $OUT/i_bench.err
timeout 600 python profiles/soak.py 1500 7003 > $OUT/i_soak.log 2>&1
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
timeout 300 python profiles/sweep_rate.py 65536 > $OUT/i_plain.log 2>&1 && \
timeout 600 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:reference_kernel -s 2 -c 1 -f -o $OUT/r02i_reference_65536 python profiles/sweep_rate.py 65536 > $OUT/i_ncu_ref.log 2>&1
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH > $OUT/i_plain2.log 2>&1 && \
timeout 900 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02i_pathtrace_c2 $BENCH > $OUT/i_ncu_full.log 2>&1
