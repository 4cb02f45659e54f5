"""The opt-in instantiation of the hierarchy kernel (sorted traversal of a large face tree + beams, RT3_BEAM_BVH=1) against the default one and
the sweep on the reference's default scene (3 288 triangles) with materials: frames and ray counts must be equal. One JSON line.
Usage: python profiles/beam_bvh_sorted_check.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi  # noqa: E402
from conftest import load_golden  # noqa: E402

g, mesh = load_golden("default_400x225")
ctx = abi.Context(0)
ctx.upload(mesh)
out = {"faces": int(mesh.n_faces)}
for w, h, spp, depth in ((96, 54, 48, 8), (64, 36, 300, 50)):
    cam = abi.reference_camera(w, h)
    p = dict(mode=abi.MODE_PATHTRACE, spp=spp, max_depth=depth, seed=11)
    sweep = ctx.render(cam, abi.make_params(w, h, **p)).copy()
    rays = ctx.stats().rays
    frames = {}
    for env in ("0", "1"):
        os.environ["RT3_BEAM_BVH"] = env
        frames[env] = ctx.render(cam, abi.make_params(w, h, flags=abi.FLAG_BVH, **p)).copy()
        st = ctx.stats()
        out[f"{w}x{h}x{spp}_beam{env}"] = {"equal_to_sweep": bool(np.array_equal(frames[env], sweep)), "rays_equal": st.rays == rays, "beam_rays": st.beam_rays,
                                          "kernel_ms": round(st.trace_kernel_ms, 3), "visits_per_ray": round(st.accel_node_visits / st.rays, 2)}
print(json.dumps(out), flush=True)
