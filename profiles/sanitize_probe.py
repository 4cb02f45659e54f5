"""A small tour of every kernel of the library for compute-sanitizer (one tool per gpurun call):
  compute-sanitizer --tool memcheck  python profiles/sanitize_probe.py
  compute-sanitizer --tool racecheck python profiles/sanitize_probe.py
Frames are tiny; only ctypes + numpy are loaded (no torch). Prints what it ran; exits non-zero on an ABI error."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi, scenes  # noqa: E402
from test_gpu_reference_mode import random_soup  # noqa: E402

ctx = abi.Context(0)
rng = np.random.default_rng(11)
w, h = 48, 27
cam = abi.reference_camera(w, h)


def with_materials(scene):
    mats = np.zeros(3, abi.MATERIAL_DTYPE)
    mats["kind"], mats["albedo"], mats["fuzz"], mats["ior"] = [0, 1, 2], [(0.7, 0.7, 0.7), (0.9, 0.8, 0.6), (1, 1, 1)], [0, 0.2, 0], [1, 1, 1.5]
    return abi.SceneArrays(faces=scene.faces, vertices=scene.vertices, face_entity=scene.face_entity, face_material=rng.integers(0, 3, scene.n_faces).astype(np.uint32),
                           spheres=scene.spheres, sphere_color=scene.sphere_color, sphere_entity=scene.sphere_entity,
                           sphere_material=rng.integers(0, 3, scene.n_spheres).astype(np.uint32), materials=mats)


def tour(name, scene, binning=("default",)):
    ctx.upload(scene)
    for flags in (0, abi.FLAG_BVH):
        ctx.render_aov(cam, abi.make_params(w, h, flags=flags))
        for b in binning if flags else ("default",):
            if b != "default":
                os.environ["RT3_BINNING"] = b
            ctx.render(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=3, max_depth=6, seed=2, flags=flags))
            os.environ.pop("RT3_BINNING", None)
    ctx.render(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=2, max_depth=6, seed=2, first_sample=3, flags=abi.FLAG_ACCUMULATE | abi.FLAG_BVH))
    ctx.read_radiance(w, h)
    print("ok", name, scene.n_faces, "faces", scene.n_spheres, "spheres", flush=True)


tour("constant-bank scene", with_materials(random_soup(rng, 300, 200)))
tour("mesh with a face tree (all three ray sorts)", with_materials(random_soup(rng, 1500, 40)), binning=("0", "1", "2"))
tour("streamed scene", with_materials(random_soup(rng, 200, 3000)))
tour("empty scene", abi.SceneArrays())
tour("cover scene", scenes.rtiow_cover(w, h)[0])
# scene assembled in device memory: a host triangle next to device-tessellated spheres
balls = [((0, 0, -3), 1.0, 12, 9, (1, 0, 0), 0), ((1.5, 0.2, -4), 0.7, 7, 5, (0, 1, 0), 1)]
nf = sum(ctx.lib.rt3_uv_sphere_faces(s[2], s[3]) for s in balls)
nv = sum(ctx.lib.rt3_uv_sphere_vertices(s[2], s[3]) for s in balls)
faces_p, verts_p = ctx.buffer_alloc(nf * 48), ctx.buffer_alloc(nv * 16)
ctx.tessellate_spheres_device(balls, 0, 0, faces_p, verts_p)
ctx.upload_device(n_faces=nf, n_vertices=nv, faces=faces_p, vertices=verts_p)
ctx.render_aov(cam, abi.make_params(w, h))
ctx.buffer_free(faces_p)
ctx.buffer_free(verts_p)
ctx.tessellate_spheres(balls, first_vertex=5)
print("ok device-built scene", flush=True)
# frame end: partition pack / unpack and the byte conversion
frame, slab, rgb = ctx.frame_alloc(w * h), ctx.buffer_alloc(w * h * 4), ctx.buffer_alloc(w * h * 4)
ctx.upload(scenes.rtiow_four_spheres(w, h)[0])
ctx.render_device(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=2, max_depth=4, tile_rows=2, part_index=1, part_count=3), frame)
ctx.pack_partition(frame, slab, w, h, 2, 1, 3)
ctx.unpack_partition(slab, frame, w, h, 2, 1, 3)
ctx.frame_bytes(frame, rgb, w, h, 3)
ctx.frame_bytes(frame, rgb, w, h, 4)
ctx.frame_read(frame, w, h)
ctx.frame_free(frame)
ctx.buffer_free(slab)
ctx.buffer_free(rgb)
print("ok frame end", flush=True)
ctx.close()
