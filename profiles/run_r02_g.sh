#!/bin/bash
# Round 2, GPU call G: randomised parity soak over the device-built scene records and the sorted traversals, the full GPU suite,
# the bench line of the final kernels, and their ncu captures (launch list; full set + FP32 op counters of the headline kernel).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python profiles/soak.py 2500 7001 > $OUT/g_soak_default.log 2>&1; echo "rc=$?" >> $OUT/g_soak_default.log
RT3_BINNING=2 timeout 500 python profiles/soak.py 800 7002 > $OUT/g_soak_warp_sorted.log 2>&1; echo "rc=$?" >> $OUT/g_soak_warp_sorted.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/g_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/g_bench.json 2> $OUT/g_bench.err; echo "bench rc=$?" >> $OUT/g_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/g_bench_reference_arm.json 2>> $OUT/g_bench.err
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
timeout 300 $BENCH > $OUT/g_plain.log 2>&1 && \
timeout 900 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02g_pathtrace_c2 $BENCH > $OUT/g_ncu_full.log 2>&1
BENCH2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH2 > $OUT/g_plain2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02g_launches.csv $BENCH2 > $OUT/g_ncu_launches.log 2>&1
ls -la $OUT > $OUT/g_listing.txt
