#!/bin/bash
# Round 2, GPU call U: ncu captures of the final build (headline kernel: full set + FP32 op counters; launch list).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH > $OUT/u_plain.log 2>&1 && \
timeout 900 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02u_pathtrace_c2 $BENCH > $OUT/u_ncu_full.log 2>&1
BENCH2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH2 > $OUT/u_plain2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02u_launches.csv $BENCH2 > $OUT/u_ncu_launches.log 2>&1
timeout 300 python profiles/bvh_c3_probe.py > $OUT/u_plain3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02u_bvh_c3 python profiles/bvh_c3_probe.py > $OUT/u_ncu_bvh.log 2>&1
