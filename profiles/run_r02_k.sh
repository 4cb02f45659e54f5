#!/bin/bash
# Round 2, GPU call K: the final build (lean level-1 bookkeeping, 16-byte loads of two pairs' -R^2 in the streamed sweep; call J ran the same
# script on a build whose drain used an inline `bfind`, which made ptxas drop the uniform loads: 198.6 ms instead of 137): smoke, full suite, soak, sweep rates, bench, and the ncu captures that profiles/ncu_counters.json is made from.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/k_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/k_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/k_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/k_pytest.log
timeout 600 python profiles/soak.py 1000 7006 > $OUT/k_soak.log 2>&1
RT3_BINNING=2 timeout 300 python profiles/soak.py 300 7007 > $OUT/k_soak_warp_sorted.log 2>&1
timeout 300 python profiles/sweep_rate.py > $OUT/k_sweep_rate.jsonl 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/k_bench.json 2> $OUT/k_bench.err; echo "bench rc=$?" >> $OUT/k_bench.err
timeout 600 python profiles/configs.py c1 c3 c5 > $OUT/k_configs.jsonl 2> $OUT/k_configs.err
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
timeout 300 python profiles/sweep_rate.py 65536 > $OUT/k_plain.log 2>&1 && \
timeout 600 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:reference_kernel -s 2 -c 1 -f -o $OUT/r02k_reference_65536 python profiles/sweep_rate.py 65536 > $OUT/k_ncu_ref.log 2>&1
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH > $OUT/k_plain2.log 2>&1 && \
timeout 900 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02k_pathtrace_c2 $BENCH > $OUT/k_ncu_full.log 2>&1
BENCH2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4"
timeout 300 $BENCH2 > $OUT/k_plain3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02k_launches.csv $BENCH2 > $OUT/k_ncu_launches.log 2>&1
