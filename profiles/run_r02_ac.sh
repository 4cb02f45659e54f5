#!/bin/bash
# Round 2, GPU call AC: ncu capture of the BEAM kernel on C2 (where do the instructions go now?)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH > $OUT/ac_plain.log 2>&1 && \
timeout 900 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02ac_pathtrace_c2_beam $BENCH > $OUT/ac_ncu_full.log 2>&1
