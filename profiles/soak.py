"""Randomised parity soak (not part of the test suite): many random scenes, scales and cameras; the sweep, the hierarchy
and the CPU oracle must agree bit for bit in reference mode (ids, t bits) and in path-traced frames.
Usage: python profiles/soak.py [n_scenes] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rt3_b200  # noqa: F401,E402
from rt3_b200 import abi  # noqa: E402
import oraclelib as ol  # noqa: E402
from test_gpu_reference_mode import random_soup  # noqa: E402
from test_gpu_stress import look_camera  # noqa: E402

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
ctx = abi.Context(0)
bad = 0
beams = 0
t0 = time.time()
for i in range(n_scenes):
    nf = int(rng.choice([0, 0, 3, 40, 500, 3000]))
    ns = int(rng.choice([0, 1, 30, 400, 700, 2500]))
    if nf + ns == 0:
        ns = 5
    scene = random_soup(rng, nf, ns)
    scale = float(10 ** rng.uniform(-2, 3))
    shift = rng.normal(0, 1, 3) * scale * float(rng.choice([0, 1, 30]))
    scene.vertices["xyz"] = (scene.vertices["xyz"].astype(np.float64) * scale + shift).astype(np.float32)
    scene.spheres[:, :3] = (scene.spheres[:, :3].astype(np.float64) * scale + shift).astype(np.float32)
    scene.spheres[:, 3] = (scene.spheres[:, 3].astype(np.float64) * scale * float(10 ** rng.uniform(-1.5, 0.5))).astype(np.float32)
    if nf:   # normals of the scaled triangles, as the reference computes them
        p = scene.vertices["xyz"].reshape(-1, 3, 3)
        a, b = p[:, 2] - p[:, 0], p[:, 1] - p[:, 0]
        c = np.stack([a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2], a[:, 2] * b[:, 0] - b[:, 2] * a[:, 0], a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]], 1).astype(np.float32)
        with np.errstate(all="ignore"):
            inv = (np.float32(1.0) / np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2], dtype=np.float32)).astype(np.float32)
        scene.faces["normal"] = c * inv[:, None]
    w, h = int(rng.integers(17, 90)), int(rng.integers(9, 60))
    eye = shift + rng.normal(0, 1, 3) * scale * 8 + np.array([0, 0, 2.0]) * scale
    target = shift + np.array([0, 0, -5.5]) * scale
    cam = look_camera(eye, target, w, h, float(rng.uniform(0.2, 1.5)))
    ctx.upload(scene)
    oframe, oprim, oent, ot = ol.oracle_reference(scene, cam, w, h)
    for flags in (0, abi.FLAG_BVH):
        frame, prim, ent, t = ctx.render_aov(cam, abi.make_params(w, h, flags=flags))
        if not (np.array_equal(prim, oprim) and np.array_equal(t.view(np.uint32), ot.view(np.uint32)) and np.array_equal(frame, oframe)):
            bad += 1
            print(f"scene {i}: reference mode differs (flags={flags}, faces={nf}, spheres={ns}, scale={scale:.3g}): {int((prim != oprim).sum())} ids", flush=True)
    if i % 4 == 0:
        mats = np.zeros(3, abi.MATERIAL_DTYPE)
        mats["kind"] = [abi.MAT_LAMBERTIAN, abi.MAT_METAL, abi.MAT_DIELECTRIC]
        mats["albedo"] = [(0.7, 0.7, 0.7), (0.9, 0.8, 0.6), (1, 1, 1)]
        mats["fuzz"] = [0, 0.1, 0]; mats["ior"] = [1, 1, 1.5]
        sc = abi.SceneArrays(faces=scene.faces, vertices=scene.vertices, face_entity=scene.face_entity, face_material=rng.integers(0, 3, nf).astype(np.uint32),
                             spheres=scene.spheres, sphere_material=rng.integers(0, 3, ns).astype(np.uint32), sphere_entity=scene.sphere_entity, materials=mats)
        ctx.upload(sc)
        pp = dict(mode=abi.MODE_PATHTRACE, spp=3, max_depth=12, seed=i)
        cpu, _, rays = ol.oracle_pathtrace(sc, cam, abi.make_params(w, h, **pp))
        for flags in (0, abi.FLAG_BVH):
            gpu = ctx.render(cam, abi.make_params(w, h, flags=flags, **pp))
            if not (np.array_equal(gpu, cpu) and ctx.stats().rays == rays):
                bad += 1
                print(f"scene {i}: path tracing differs (flags={flags}, faces={nf}, spheres={ns}, scale={scale:.3g}): {int((gpu != cpu).sum())} pixels", flush=True)
        # one row at many samples per pixel: a chunk of path items then lies in a few pixels, and the hierarchy kernels trace its primary
        # rays against the candidates a beam's walk of the trees collected (rt3_kernels.cuh, beam_for_chunk_bvh); sweep kernels of
        # resident sphere scenes do the same with their own candidate lists
        y = int(rng.integers(0, h))
        pb = dict(mode=abi.MODE_PATHTRACE, spp=int(rng.choice([24, 48, 128, 300])), max_depth=6, seed=i, tile_rows=1, part_index=y, part_count=h)
        cpu, _, rays = ol.oracle_pathtrace(sc, cam, abi.make_params(w, h, **pb))
        for flags in (0, abi.FLAG_BVH):
            gpu = ctx.render(cam, abi.make_params(w, h, flags=flags, **pb))
            st = ctx.stats()
            beams += st.beam_rays
            if not (np.array_equal(gpu[y], cpu[y]) and st.rays == rays):
                bad += 1
                print(f"scene {i}: path tracing of row {y} at {pb['spp']} spp differs (flags={flags}, faces={nf}, spheres={ns}, scale={scale:.3g}): {int((gpu[y] != cpu[y]).sum())} pixels", flush=True)
print(f"{n_scenes} scenes, {bad} mismatches, {beams} primary rays through candidate lists, {time.time() - t0:.1f} s")
