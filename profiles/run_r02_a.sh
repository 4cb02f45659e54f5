#!/bin/bash
# Round 2, GPU call A: parity tests (incl. the full-shape bands and the independent pin), bench (with the C4 leg),
# survivor statistics, upload times, then the ncu captures (launch list; full set + FP32 op counters of the headline kernel).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/a_smi.txt 2>&1
nproc > $OUT/a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/a_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/a_bench.json 2> $OUT/a_bench.err; echo "bench rc=$?" >> $OUT/a_bench.err
timeout 300 python profiles/survivors.py 32 > $OUT/a_survivors.json 2>&1
timeout 300 python profiles/upload_time.py > $OUT/a_upload_time.txt 2>&1
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__sass_thread_inst_executed_ops_fadd_fmul_ffma_pred_on.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,smsp__inst_executed_op_ffma2.sum"
timeout 300 $BENCH > $OUT/a_plain.log 2>&1 && \
timeout 900 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02a_pathtrace_c2 $BENCH > $OUT/a_ncu_full.log 2>&1
BENCH2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH2 > $OUT/a_plain2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02a_launches.csv $BENCH2 > $OUT/a_ncu_launches.log 2>&1
timeout 300 python profiles/sweep_rate.py 65536 > $OUT/a_sweep_65536.json 2>&1 && \
timeout 600 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:reference_kernel -s 2 -c 1 -f -o $OUT/r02a_reference_65536 python profiles/sweep_rate.py 65536 > $OUT/a_ncu_ref.log 2>&1
ls -la $OUT > $OUT/a_listing.txt
