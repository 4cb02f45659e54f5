"""Times one build of the library (RT3_CORE_LIB) on BASELINE C2 (sweep and hierarchy) and, with --c3, on C3 through the hierarchy.
Prints one JSON line with kernel times and a frame checksum (all variants must produce the same frames).
Usage: RT3_CORE_LIB=<lib.so> python profiles/variants.py <label> [--c3 [--spp N] | --c2bvh | --c5]"""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi, scenes  # noqa: E402

label = sys.argv[1]
ctx = abi.Context(0)
out = {"variant": label, "lib": os.path.basename(abi.CORE_LIB_PATH), "binning": os.environ.get("RT3_BINNING", "default")}


def timed(cam, params, reps=3):
    best, frame = None, None
    for _ in range(reps):
        frame = ctx.render(cam, params)
        st = ctx.stats()
        best = st.trace_kernel_ms if best is None else min(best, st.trace_kernel_ms)
    return round(best, 3), zlib.crc32(frame.tobytes()), st


if "--c3" in sys.argv:
    import fullsize
    spp = int(sys.argv[sys.argv.index("--spp") + 1]) if "--spp" in sys.argv else 256
    scene, cam = fullsize.c3_scene(1920, 1080)
    ctx.upload(scene)
    ms, crc, st = timed(cam, abi.make_params(1920, 1080, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=50, seed=1, flags=abi.FLAG_BVH), reps=2)
    out.update(c3_spp=spp, c3_bvh_kernel_ms=ms, c3_crc=crc, c3_grays_s=round(st.rays / ms / 1e6, 3), c3_visits_per_ray=round(st.accel_node_visits / st.rays, 2), c3_visits=st.accel_node_visits, c3_tests=st.accel_prim_tests)
elif "--c2bvh" in sys.argv:
    scene, cam = scenes.rtiow_cover(1200, 800)
    ctx.upload(scene)
    ms, crc, st = timed(cam, abi.make_params(1200, 800, mode=abi.MODE_PATHTRACE, spp=500, max_depth=50, seed=1, tile_rows=2, flags=abi.FLAG_BVH), reps=2)
    out.update(c2_bvh_kernel_ms=ms, c2_crc=crc, c2_grays_s=round(st.rays / ms / 1e6, 3), c2_visits_per_ray=round(st.accel_node_visits / st.rays, 2), c2_visits=st.accel_node_visits, c2_tests=st.accel_prim_tests)
elif "--c5" in sys.argv:
    scene, cam = scenes.random_spheres(1000000)
    ctx.upload(scene)
    ms, crc, st = timed(cam, abi.make_params(1920, 1080, mode=abi.MODE_PATHTRACE, spp=64, max_depth=1, seed=1, flags=abi.FLAG_BVH | abi.FLAG_NO_JITTER), reps=2)
    out.update(c5_bvh_kernel_ms=ms, c5_crc=crc, c5_grays_s=round(st.rays / ms / 1e6, 3), c5_visits_per_ray=round(st.accel_node_visits / st.rays, 2), c5_visits=st.accel_node_visits, c5_tests=st.accel_prim_tests)
else:
    scene, cam = scenes.rtiow_cover(1200, 800)
    ctx.upload(scene)
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    ms, crc, st = timed(cam, abi.make_params(1200, 800, mode=abi.MODE_PATHTRACE, spp=500, max_depth=50, seed=1, tile_rows=2), reps=reps)
    out.update(c2_sweep_kernel_ms=ms, c2_crc=crc, c2_grays_s=round(st.rays / ms / 1e6, 3), beam=os.environ.get("RT3_BEAM", "default"), rays=st.rays,
               beam_rays=st.beam_rays, beam_tests_per_ray=round(st.beam_tests / max(st.beam_rays, 1), 2))
    if "--eighth" in sys.argv:   # one rank's share of the frame at 8 GPUs: where the tail of the kernel shows
        ms8, _, st8 = timed(cam, abi.make_params(1200, 800, mode=abi.MODE_PATHTRACE, spp=500, max_depth=50, seed=1, tile_rows=1, part_index=3, part_count=8), reps=reps)
        out.update(c2_eighth_kernel_ms=ms8, c2_eighth_ideal_ms=round(ms * st8.rays / st.rays, 3))
if "--c3" in sys.argv or "--c2bvh" in sys.argv or "--c5" in sys.argv:   # hierarchy kernels: primary rays through the candidates of a beam's walk (RT3_BEAM_BVH=0: off)
    out.update(beam_bvh=os.environ.get("RT3_BEAM_BVH", "default"), rays=st.rays, beam_rays=st.beam_rays, beam_tests_per_ray=round(st.beam_tests / max(st.beam_rays, 1), 2),
               tests_per_ray=round(st.accel_prim_tests / st.rays, 2))
print(json.dumps(out), flush=True)
