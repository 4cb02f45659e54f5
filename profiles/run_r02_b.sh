#!/bin/bash
# Round 2, GPU call B: validates the device-side scene build (rt3_upload.cuh) and the four-ray reference kernel, measures the
# prerender cost, the hierarchy statistics of C3 / C5, and captures the intersection kernel (reference_kernel, 65 536 spheres).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/b_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/b_pytest.log
timeout 300 python profiles/upload_time.py > $OUT/b_upload_time.jsonl 2> $OUT/b_upload_time.err
timeout 300 python profiles/sweep_rate.py 484 4096 65536 1048576 > $OUT/b_sweep_rate_r4.jsonl 2>&1
RT3_CORE_LIB=$PWD/profiles/librt3cuda_ref2.so timeout 300 python profiles/sweep_rate.py 484 4096 65536 1048576 > $OUT/b_sweep_rate_r2.jsonl 2>&1
timeout 900 python profiles/configs.py c1 c2 c3 c5 --oracle > $OUT/b_configs.jsonl 2> $OUT/b_configs.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-c4 > $OUT/b_bench.json 2> $OUT/b_bench.err
FP="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"
timeout 300 python profiles/sweep_rate.py 65536 > $OUT/b_plain.log 2>&1 && \
timeout 600 ncu --set full --metrics $FP --clock-control none --import-source on -k regex:reference_kernel -s 2 -c 1 -f -o $OUT/r02b_reference_65536 python profiles/sweep_rate.py 65536 > $OUT/b_ncu_ref.log 2>&1
ls -la $OUT > $OUT/b_listing.txt
