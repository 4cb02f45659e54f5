"""Intersection-loop roofline sweep (BASELINE config C5 shape): reference-mode render of N random
analytic spheres at 1920x1080, one primary ray per pixel, brute force. Prints the algorithmic
FP32 rate (17 FLOP per ray-sphere test) of reference_kernel alone (CUDA events inside the library).
Usage: python profiles/sweep_rate.py [N ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi, scenes  # noqa: E402

W, H = 1920, 1080
ns = [int(a) for a in sys.argv[1:]] or [1024, 2048, 16384, 131072, 1048576]
ctx = abi.Context(0)
peak = ctx.measure_fma_peak()
for n in ns:
    scene, cam = scenes.random_spheres(n, width=W, height=H)
    ctx.upload(scene)
    params = abi.make_params(W, H, mode=abi.MODE_REFERENCE)
    best = None
    for _ in range(3 if n <= 131072 else 1):
        ctx.render(cam, params)
        st = ctx.stats()
        best = st.trace_kernel_ms if best is None else min(best, st.trace_kernel_ms)
    tests = W * H * n
    tflops = 17 * tests / (best * 1e-3) / 1e12
    cyc = best * 1e-3 * 1.965e9 * 148 * 4 / (tests / 64)  # SMSP cycles per (primitive x ray pair x warp)
    print(json.dumps({"spheres": n, "kernel_ms": round(best, 3), "mrays_s": round(W * H / best / 1e3, 2), "algorithmic_tflops": round(tflops, 2),
                      "frac_of_measured_ffma_peak": round(tflops / peak, 3), "smsp_cycles_per_prim_pair": round(cyc, 2), "ffma_peak_tflops": round(peak, 2)}))
