#!/bin/bash
# Round 2, GPU call O: compute-sanitizer is closed on this pool (call N: "closed and stays closed"), so the bounds are checked by the code
# itself: the RT3_DEBUG_ASSERTS build (device-side asserts at every data-dependent index) runs the kernel tour, the soak and the GPU suite.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
export RT3_CORE_LIB=$PWD/profiles/asserts/librt3cuda.so   # nvcc <the flags of __graft_entry__.build()> -DRT3_DEBUG_ASSERTS -o profiles/asserts/librt3cuda.so raytracer-3_b200/csrc/rt3_core.cu
timeout 300 python profiles/sanitize_probe.py > $OUT/o_probe_asserts.log 2>&1; echo "rc=$?" >> $OUT/o_probe_asserts.log
timeout 600 python profiles/soak.py 1500 7010 > $OUT/o_soak_asserts.log 2>&1; echo "rc=$?" >> $OUT/o_soak_asserts.log
RT3_BINNING=2 timeout 300 python profiles/soak.py 400 7011 > $OUT/o_soak_asserts_b2.log 2>&1; echo "rc=$?" >> $OUT/o_soak_asserts_b2.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/o_pytest_asserts.log 2>&1; echo "pytest rc=$?" >> $OUT/o_pytest_asserts.log
unset RT3_CORE_LIB
timeout 900 python -m pytest tests -m gpu -x -q -k "bvh or hierarchy or full_size or scene_construction" > $OUT/o_pytest_release.log 2>&1; echo "pytest rc=$?" >> $OUT/o_pytest_release.log
