"""The five BASELINE.json configs on one B200 (C4: this GPU's share is the whole frame; the multi-GPU split is
bench.py --gpus N). Each config is rendered through the brute-force sweep and through the hierarchy
(RT3_FLAG_BVH) -- the frames must be identical -- at the configured resolution and depth; spp is reduced
where the brute-force side would take minutes, and says so. Device time is rt3_stats.device_ms
(CUDA events around clear + trace + resolve). Usage: python profiles/configs.py [c1 c2 c3 c4 c5] [--oracle]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi, scenes  # noqa: E402


import fullsize  # noqa: E402  (tests/fullsize.py: the configurations' real shapes and the oracle band check)

c3_scene = fullsize.c3_scene


CONFIGS = {
    "c1": dict(what="RTIOW 4-sphere scene, 400x225, 100 spp, depth 50", w=400, h=225, spp=100, depth=50, full_spp=100,
               scene=lambda w, h: scenes.rtiow_four_spheres(w, h)),
    "c2": dict(what="RTIOW cover scene (484 spheres), 1200x800, 500 spp, depth 50", w=1200, h=800, spp=500, depth=50, full_spp=500,
               scene=lambda w, h: scenes.rtiow_cover(w, h)),
    "c3": dict(what="100 350 triangles + 3 spheres, 1920x1080, depth 50", w=1920, h=1080, spp=4, depth=50, full_spp=256, scene=c3_scene),
    "c4": dict(what="cover scene at 3840x2160, depth 50 (one GPU renders the whole frame here)", w=3840, h=2160, spp=64, depth=50, full_spp=1024,
               scene=lambda w, h: scenes.rtiow_cover(w, h)),
    "c5": dict(what="10^6 random spheres, 1920x1080, depth 1, no jitter", w=1920, h=1080, spp=2, depth=1, full_spp=64,
               scene=lambda w, h: scenes.random_spheres(1000000, width=w, height=h), flags=abi.FLAG_NO_JITTER),
}

ctx = abi.Context(0)
for name in ([a for a in sys.argv[1:] if not a.startswith("--")] or list(CONFIGS)):
    c = CONFIGS[name]
    scene, cam = c["scene"](c["w"], c["h"])
    ctx.upload(scene)
    up = ctx.stats()
    out = {"config": name, "upload_ms": round(up.upload_ms, 3), "upload_h2d_ms": round(up.h2d_ms, 3), "upload_build_kernels_ms": round(up.upload_device_ms, 3), "what": c["what"], "spp": c["spp"], "spp_in_BASELINE": c["full_spp"], "faces": scene.n_faces, "spheres": scene.n_spheres}
    frames = {}
    for path, flag in (("sweep", 0), ("bvh", abi.FLAG_BVH)):
        params = abi.make_params(c["w"], c["h"], mode=abi.MODE_PATHTRACE, spp=c["spp"], max_depth=c["depth"], seed=1, flags=c.get("flags", 0) | flag)
        ctx.render(cam, params)                      # warm-up (and hierarchy build)
        frames[path] = ctx.render(cam, params)
        st = ctx.stats()
        out[path + "_device_ms"] = round(st.device_ms, 3)
        out[path + "_mrays_s"] = round(st.rays / st.device_ms / 1e3, 1)
        out["rays"] = st.rays
        if flag:
            out["bvh_build_ms"] = round(st.accel_build_ms, 3)
            out["bvh_node_visits_per_ray"] = round(st.accel_node_visits / max(st.rays, 1), 2)
            out["bvh_prim_tests_per_ray"] = round(st.accel_prim_tests / max(st.rays, 1), 2)
    out["frames_identical"] = bool(np.array_equal(frames["sweep"], frames["bvh"]))
    if "--oracle" in sys.argv:
        # row bands of the full-size frame (and sample windows where a row is minutes of brute force) against the CPU oracle
        fc = fullsize.CONFIGS[name]
        cpu = fullsize.oracle_bands(fc, scene, cam)
        for path, flag in (("sweep", 0), ("bvh", abi.FLAG_BVH)):
            if name == "c5" and path == "sweep":
                continue
            diff, rays_equal = fullsize.bands_match(fullsize.gpu_bands(ctx, fc, cam, flag), cpu)
            out["matches_oracle_bands_" + path] = bool(diff == 0 and rays_equal)
        out["oracle_bands"] = [list(b) for b in fullsize.bands(fc)]
    if c["spp"] != c["full_spp"]:
        # the configured sample count through the hierarchy (and through the sweep where that takes seconds, not minutes)
        for path, flag in (("bvh", abi.FLAG_BVH),) + ((("sweep", 0),) if name == "c4" else ()):
            params = abi.make_params(c["w"], c["h"], mode=abi.MODE_PATHTRACE, spp=c["full_spp"], max_depth=c["depth"], seed=1, flags=c.get("flags", 0) | flag)
            ctx.render(cam, params)
            st = ctx.stats()
            out[path + "_full_spp_device_ms"] = round(st.device_ms, 2)
            out[path + "_full_spp_mrays_s"] = round(st.rays / st.device_ms / 1e3, 1)
            out["full_spp_rays"] = st.rays
    print(json.dumps(out), flush=True)
