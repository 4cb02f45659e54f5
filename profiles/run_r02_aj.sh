#!/bin/bash
# Round 2, GPU call AJ: the final tree as the driver will see it -- GPU suite, smoke, both bench arms -- then soak, the A/B of the regeneration
# batches of the hierarchy's beams, the ncu captures of the headline kernel (full set + FP32 op counters; launch list) and the five BASELINE configs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/aj_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/aj_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/aj_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/aj_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/aj_bench.json 2> $OUT/aj_bench.err; echo "bench rc=$?" >> $OUT/aj_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/aj_bench_reference.json 2> $OUT/aj_bench_reference.err; echo "rc=$?" >> $OUT/aj_bench_reference.err
timeout 300 python profiles/soak.py 2000 31337 > $OUT/aj_soak.log 2>&1; echo "rc=$?" >> $OUT/aj_soak.log
: > $OUT/aj_variants.jsonl
timeout 200 python profiles/variants.py abatches-16 --c5 >> $OUT/aj_variants.jsonl 2>> $OUT/aj_variants.err
RT3_CORE_LIB=$PWD/profiles/librt3cuda_abatches4.so timeout 200 python profiles/variants.py abatches-4 --c5 >> $OUT/aj_variants.jsonl 2>> $OUT/aj_variants.err
timeout 200 python profiles/variants.py abatches-16 --c2bvh >> $OUT/aj_variants.jsonl 2>> $OUT/aj_variants.err
RT3_CORE_LIB=$PWD/profiles/librt3cuda_abatches4.so timeout 200 python profiles/variants.py abatches-4 --c2bvh >> $OUT/aj_variants.jsonl 2>> $OUT/aj_variants.err
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH > $OUT/aj_plain.log 2>&1 && \
timeout 600 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02aj_pathtrace_c2 $BENCH > $OUT/aj_ncu_full.log 2>&1
BENCH2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH2 > $OUT/aj_plain2.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02aj_launches.csv $BENCH2 > $OUT/aj_ncu_launches.log 2>&1
timeout 500 python profiles/configs.py --oracle > $OUT/aj_configs.jsonl 2> $OUT/aj_configs.err; echo "rc=$?" >> $OUT/aj_configs.err
tail -3 $OUT/aj_pytest.log; tail -2 $OUT/aj_smoke.log; cut -c1-200 $OUT/aj_bench.json; tail -1 $OUT/aj_soak.log; cut -c1-200 $OUT/aj_variants.jsonl
