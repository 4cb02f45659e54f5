#!/bin/bash
# Round 2, GPU call AF: ncu capture of the BEAM headline kernel (pathtrace_kernel<1,1,0,0,1>, BASELINE C2): full set + FP32 op counters
# + source page, taken only after the same command exited 0 without ncu. (Call AE's capture was lost with its session.)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH > $OUT/af_plain.log 2>&1 && \
timeout 900 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02af_pathtrace_c2 $BENCH > $OUT/af_ncu_full.log 2>&1
echo "ncu rc=$?" >> $OUT/af_ncu_full.log
ls -la $OUT
