import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import rt3_b200
from rt3_b200 import abi
import importlib.util
spec = importlib.util.spec_from_file_location("cfg", "/root/repo/profiles/configs.py")
sys.argv = ["x", "none"]
try:
    cfg = importlib.util.module_from_spec(spec); spec.loader.exec_module(cfg)
except KeyError:
    pass
scene, cam = cfg.c3_scene(1920, 1080)
ctx = abi.Context(0); ctx.upload(scene)
for spp in (1, 4, 16, 64):
    p = abi.make_params(1920, 1080, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=50, seed=1, flags=abi.FLAG_BVH)
    ctx.render(cam, p); ctx.render(cam, p); st = ctx.stats()
    print(spp, "spp", round(st.device_ms, 2), "ms", round(st.rays / st.device_ms / 1e3, 1), "Mrays/s", "visits/ray", round(st.accel_node_visits / st.rays, 1), "tests/ray", round(st.accel_prim_tests / st.rays, 2), "rays/path", round(st.rays / (1920 * 1080 * spp), 2))
for depth in (1, 2, 5, 10, 50):
    p = abi.make_params(1920, 1080, mode=abi.MODE_PATHTRACE, spp=1, max_depth=depth, seed=1, flags=abi.FLAG_BVH)
    ctx.render(cam, p); ctx.render(cam, p); st = ctx.stats()
    print("depth", depth, round(st.device_ms, 2), "ms; trace kernel", round(st.trace_kernel_ms, 2), "rays", st.rays)
p = abi.make_params(1920, 1080, flags=abi.FLAG_BVH)
ctx.render(cam, p); ctx.render(cam, p); st = ctx.stats()
print("reference mode", round(st.device_ms, 2), "ms")
p = abi.make_params(1920, 1080, mode=abi.MODE_PATHTRACE, spp=256, max_depth=50, seed=1, flags=abi.FLAG_BVH)
ctx.render(cam, p); st = ctx.stats()
print("C3 at the configured 256 spp:", round(st.device_ms, 1), "ms", round(st.rays / st.device_ms / 1e3, 1), "Mrays/s", st.rays, "rays")
