#!/bin/bash
# Round 2, GPU call AA: the final tree as the driver will see it -- GPU suite, smoke, both bench arms -- plus the ncu capture of the hierarchy
# traversal on the C3 mesh with the default three-class ray sort (the r02u capture was taken with two classes) and a soak run.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/aa_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/aa_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/aa_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/aa_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/aa_bench_reference.json 2> $OUT/aa_bench_reference.err; echo "rc=$?" >> $OUT/aa_bench_reference.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/aa_bench.json 2> $OUT/aa_bench.err; echo "bench rc=$?" >> $OUT/aa_bench.err
timeout 300 python profiles/soak.py 1500 7031 > $OUT/aa_soak.log 2>&1; echo "rc=$?" >> $OUT/aa_soak.log
timeout 300 python profiles/bvh_c3_probe.py > $OUT/aa_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pathtrace_kernel -s 1 -c 1 -f -o $OUT/r02aa_bvh_c3 python profiles/bvh_c3_probe.py > $OUT/aa_ncu_bvh.log 2>&1
