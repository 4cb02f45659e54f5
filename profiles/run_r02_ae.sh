#!/bin/bash
# Round 2, GPU call AE: the BEAM build (4 regeneration batches) as the driver will see it -- GPU suite, smoke, soak, both bench arms --
# and its ncu captures (headline kernel: full set + FP32 op counters; launch list).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/ae_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ae_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/ae_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/ae_smoke.log
timeout 400 python profiles/soak.py 3000 7051 > $OUT/ae_soak.log 2>&1; echo "rc=$?" >> $OUT/ae_soak.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/ae_bench.json 2> $OUT/ae_bench.err; echo "bench rc=$?" >> $OUT/ae_bench.err
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH > $OUT/ae_plain.log 2>&1 && \
timeout 900 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02ae_pathtrace_c2 $BENCH > $OUT/ae_ncu_full.log 2>&1
BENCH2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH2 > $OUT/ae_plain2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02ae_launches.csv $BENCH2 > $OUT/ae_ncu_launches.log 2>&1
