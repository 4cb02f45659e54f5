"""Prerender cost (the reference's contract is prerender once, then render: src/Main.cpp:284-285): rt3_scene_upload for
484 / 10^5 / 10^6 spheres and for the 100 350-triangle C3 mesh, and the first render through the hierarchy (which builds it).
upload_ms: wall clock of the call; h2d_ms: the copies of the input arrays; build_kernels_ms: the kernels that derive bounds, boxes,
prefilter records and the basis (rt3_upload.cuh); device_arrays: the same scene when it already lies in device memory
(rt3_scene_upload_device). Usage: python profiles/upload_time.py   (one JSON line per scene)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi, scenes  # noqa: E402

ctx = abi.Context(0)
ctx.upload(scenes.random_spheres(16, width=64, height=36)[0])   # first call: module load, context warm-up


def measure(name, scene, cam):
    out = {"scene": name, "faces": scene.n_faces, "spheres": scene.n_spheres}
    for rep in range(2):
        t0 = time.perf_counter()
        ctx.upload(scene)
        t1 = time.perf_counter()
        st = ctx.stats()
        out.update(upload_ms=round(st.upload_ms, 3), upload_wall_py_ms=round((t1 - t0) * 1e3, 3), h2d_ms=round(st.h2d_ms, 3), build_kernels_ms=round(st.upload_device_ms, 3))
    t1 = time.perf_counter()
    ctx.render(cam, abi.make_params(64, 36, flags=abi.FLAG_BVH))
    out["first_bvh_render_ms"] = round((time.perf_counter() - t1) * 1e3, 3)
    out["bvh_build_ms"] = round(ctx.stats().accel_build_ms, 3)
    names = ("faces", "vertices", "face_material", "face_entity", "spheres", "sphere_color", "sphere_material", "sphere_entity", "materials")
    ptr = {n: ctx.to_device(getattr(scene, n)) for n in names}
    for rep in range(2):
        ctx.upload_device(n_faces=scene.n_faces, n_vertices=len(scene.vertices), n_spheres=scene.n_spheres, n_materials=len(scene.materials), **ptr)
        st = ctx.stats()
        out["device_arrays"] = {"upload_ms": round(st.upload_ms, 3), "build_kernels_ms": round(st.upload_device_ms, 3)}
    for p in ptr.values():
        if p:
            ctx.buffer_free(p)
    print(json.dumps(out), flush=True)


for n in (484, 100000, 1000000):
    measure(f"{n} random spheres", *scenes.random_spheres(n, width=64, height=36))
import fullsize  # noqa: E402
measure("C3 mesh (100 350 triangles + 3 spheres)", *fullsize.c3_scene(64, 36))
