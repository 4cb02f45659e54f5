import sys, time
sys.path.insert(0, "/root/repo")
import rt3_b200
from rt3_b200 import abi, scenes
ctx = abi.Context(0)
for n in (484, 100000, 1000000):
    scene, cam = scenes.random_spheres(n, width=64, height=36)
    t0 = time.perf_counter(); ctx.upload(scene); t1 = time.perf_counter()
    ctx.render(cam, abi.make_params(64, 36, flags=abi.FLAG_BVH)); t2 = time.perf_counter()
    print(n, "upload ms", round((t1 - t0) * 1e3, 2), "first bvh render ms", round((t2 - t1) * 1e3, 2), "build ms", round(ctx.stats().accel_build_ms, 3))
