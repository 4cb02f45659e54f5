#!/bin/bash
# Round 2, GPU call AL: the final tree (call AJ's plus RT3_BEAM_MIN_BATCH = 8) as the driver will see it -- GPU suite, smoke, both bench arms, soak --
# and the ncu captures of its headline kernel (full set + FP32 op counters; launch list).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/al_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/al_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/al_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/al_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/al_bench.json 2> $OUT/al_bench.err; echo "bench rc=$?" >> $OUT/al_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/al_bench_reference.json 2> $OUT/al_bench_reference.err; echo "rc=$?" >> $OUT/al_bench_reference.err
timeout 200 python profiles/soak.py 1000 777 > $OUT/al_soak.log 2>&1; echo "rc=$?" >> $OUT/al_soak.log
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH > $OUT/al_plain.log 2>&1 && \
timeout 600 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02al_pathtrace_c2 $BENCH > $OUT/al_ncu_full.log 2>&1
BENCH2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH2 > $OUT/al_plain2.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02al_launches.csv $BENCH2 > $OUT/al_ncu_launches.log 2>&1
tail -3 $OUT/al_pytest.log; tail -2 $OUT/al_smoke.log; cut -c1-200 $OUT/al_bench.json; tail -2 $OUT/al_soak.log
