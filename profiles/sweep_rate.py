"""Intersection-loop roofline sweep (BASELINE config C5 shape): reference-mode
render of N random analytic spheres at 1920x1080, one un-jittered primary ray per pixel.
For every N: the brute-force sweep (reference_kernel; algorithmic FP32 rate at 17 FLOP per
ray-sphere test, SMSP cycles per test) and the same frame through the hierarchy (RT3_FLAG_BVH),
with a check that the two frames are identical. Kernel times are CUDA events inside the library.
Usage: python profiles/sweep_rate.py [N ...]   (one JSON line per N)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi, scenes  # noqa: E402

W, H = 1920, 1080
ns = [int(a) for a in sys.argv[1:]] or [1 << k for k in range(10, 21, 2)]
ctx = abi.Context(0)
peak = ctx.measure_fma_peak()
for n in ns:
    scene, cam = scenes.random_spheres(n, width=W, height=H)
    ctx.upload(scene)
    out = {"spheres": n}
    frames = {}
    for name, flags in (("sweep", 0), ("bvh", abi.FLAG_BVH)):
        params = abi.make_params(W, H, mode=abi.MODE_REFERENCE, flags=flags)
        best = None
        for _ in range(3 if n <= 131072 or flags else 1):
            frames[name] = ctx.render(cam, params)
            st = ctx.stats()
            best = st.trace_kernel_ms if best is None else min(best, st.trace_kernel_ms)
        out[name + "_kernel_ms"] = round(best, 3)
        out[name + "_mrays_s"] = round(W * H / best / 1e3, 2)
        if flags:
            out["bvh_build_ms"] = round(st.accel_build_ms, 3)
            out["bvh_node_visits_per_ray"] = round(st.accel_node_visits / st.rays, 1)
            out["bvh_prim_tests_per_ray"] = round(st.accel_prim_tests / st.rays, 2)
        else:
            tests = W * H * n
            tflops = 17 * tests / (best * 1e-3) / 1e12
            out["sweep_algorithmic_tflops"] = round(tflops, 2)
            out["sweep_frac_of_measured_ffma_peak"] = round(tflops / peak, 3)
            out["sweep_smsp_cycles_per_warp_test"] = round(best * 1e-3 * 1.965e9 * 148 * 4 / (tests / 32), 2)
    out["frames_identical"] = bool(np.array_equal(frames["sweep"], frames["bvh"]))
    out["ffma_peak_tflops"] = round(peak, 2)
    print(json.dumps(out), flush=True)
