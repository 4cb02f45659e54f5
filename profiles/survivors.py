"""Level-1 survivors per ray (VERDICT r01 "what's weak": the number that decides what to cut next in pathtrace_kernel).
Runs the RT3_SURVIVOR_STATS debug build of the library (profiles/librt3cuda_survivors.so, built by
`nvcc ... -DRT3_SURVIVOR_STATS -o profiles/librt3cuda_survivors.so raytracer-3_b200/csrc/rt3_core.cu`) on the cover scene.
Usage: python profiles/survivors.py [spp]     (one JSON line)"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["RT3_CORE_LIB"] = os.path.join(ROOT, "profiles", "librt3cuda_survivors.so")
sys.path.insert(0, ROOT)
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi, scenes  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H = 1200, 800
ctx = abi.Context(0)
scene, cam = scenes.rtiow_cover(W, H)
ctx.upload(scene)
ctx.lib.rt3_debug_survivor_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
out = (C.c_uint64 * 4)()
params = abi.make_params(W, H, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=50, seed=1)
ctx.render(cam, params)
ctx.lib.rt3_debug_survivor_stats(ctx.handle, out)   # clears the warm-up's counts
ctx.render(cam, params)
st = ctx.stats()
assert ctx.lib.rt3_debug_survivor_stats(ctx.handle, out) == 0
pairs, warp_iters, warp_drains, live_lanes = (int(v) for v in out)
print(json.dumps({"scene": "cover (C2)", "spp": spp, "rays": st.rays, "n_spheres": scene.n_spheres,
                  "survivors_per_ray": pairs / max(live_lanes, 1), "survivor_fraction": pairs / max(live_lanes, 1) / scene.n_spheres,
                  "drain_iterations_per_warp_drain": warp_iters / max(warp_drains, 1),
                  "live_lanes_per_warp_drain": live_lanes / max(warp_drains, 1),
                  "lane_utilisation_in_drain": pairs / max(warp_iters * 32, 1)}))
