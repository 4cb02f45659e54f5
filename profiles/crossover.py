import json, os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import rt3_b200
from rt3_b200 import abi, scenes
W, H = 960, 540
ctx = abi.Context(0)
for n in [int(a) for a in sys.argv[1:]]:
    scene, cam = scenes.random_spheres(n, width=W, height=H)
    ctx.upload(scene)
    out = {"n": n}
    for name, params in (("ref", abi.make_params(W, H, mode=abi.MODE_REFERENCE)), ("path", abi.make_params(W, H, mode=abi.MODE_PATHTRACE, spp=8, max_depth=8, seed=1))):
        best = None
        for _ in range(3):
            ctx.render(cam, params); st = ctx.stats()
            best = st.trace_kernel_ms if best is None else min(best, st.trace_kernel_ms)
        out[name + "_ms"] = round(best, 3)
        out[name + "_cyc_per_test"] = round(best * 1e-3 * 1.965e9 * 592 / (st.rays * n / 32), 2)
    print(json.dumps(out), flush=True)
