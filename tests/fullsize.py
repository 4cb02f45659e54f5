"""BASELINE.json's five configurations at their REAL shapes, and the row-band comparison against the CPU oracle.

The oracle is brute force on the host, so a full C2..C5 frame would take hours; a row band of the full-size frame does
not: `rt3_params.tile_rows = 1, part_count = height, part_index = y` renders exactly row y of the full frame (same
pixels, same per-pixel RNG counters -- results do not depend on the partition), on the GPU and in the oracle alike
(reference loop: src/lib/renderer/SequentialRenderer.cpp:269-308). Where even one row at the configured sample count
is minutes of brute force (C3: 100 350 faces per ray segment; C5: 10^6 spheres), the band is a row AND a window of
the samples: `first_sample = s, spp = n` renders samples [s, s + n) of the configured 256 / 64 -- the same paths with
the same counters, whose integer sums the full frame adds up (RT3_FLAG_ACCUMULATE, tested bit for bit elsewhere).
Bands are sized for <= 60 s of oracle time per configuration on 16 host threads. Used by tests/test_gpu_full_size.py and profiles/configs.py (`matches_oracle_bands`)."""
import numpy as np

import oraclelib as ol
from rt3_b200 import abi, scenes


def c3_scene(w, h):
    """~100k triangles (create_sphere(.., 225, 225, ..), reference Sphere.cpp tessellation: 100 350 faces) + 3 analytic spheres."""
    import hostlib
    hs = hostlib.HostScene()
    hs.add_sphere((0, 0, -3), 1.0, 225, 225, (0.8, 0.3, 0.3))
    mesh = hs.flatten()
    mats = np.zeros(3, abi.MATERIAL_DTYPE)
    mats["kind"] = [abi.MAT_LAMBERTIAN, abi.MAT_METAL, abi.MAT_DIELECTRIC]
    mats["albedo"] = [(0.8, 0.8, 0.0), (0.8, 0.6, 0.2), (1, 1, 1)]
    mats["fuzz"] = [0, 0.1, 0]
    mats["ior"] = [1, 1, 1.5]
    spheres = np.array([(0, -101, -3, 100), (2.2, 0, -3, 1), (-2.2, 0, -3, 1)], np.float32)
    scene = abi.SceneArrays(faces=mesh.faces, vertices=mesh.vertices, face_entity=mesh.face_entity, spheres=spheres,
                            sphere_material=np.arange(3, dtype=np.uint32), sphere_entity=np.arange(1, 4, dtype=np.uint32), materials=mats)
    return scene, abi.reference_camera(w, h)


CONFIGS = {
    # rows: the bands compared with the oracle, as (first row, row count); windows: sample windows (first_sample, spp)
    # rendered per band row, default the whole configured range
    "c1": dict(what="RTIOW 4-sphere scene, 400x225, 100 spp, depth 50", w=400, h=225, spp=100, depth=50, flags=0,
               scene=scenes.rtiow_four_spheres, rows=[(0, 225)]),
    "c2": dict(what="RTIOW cover scene (484 spheres), 1200x800, 500 spp, depth 50", w=1200, h=800, spp=500, depth=50, flags=0,
               scene=scenes.rtiow_cover, rows=[(96, 4), (400, 4), (640, 4), (796, 4)]),
    "c3": dict(what="100 350 triangles + 3 spheres, 1920x1080, 256 spp, depth 50", w=1920, h=1080, spp=256, depth=50, flags=0,
               scene=c3_scene, rows=[(470, 1), (760, 1)], windows=[(0, 2), (254, 2)]),
    "c4": dict(what="cover scene at 3840x2160, 1024 spp, depth 50", w=3840, h=2160, spp=1024, depth=50, flags=0,
               scene=scenes.rtiow_cover, rows=[(1100, 1), (1900, 1)]),
    "c5": dict(what="10^6 random spheres, 1920x1080, 64 spp, depth 1, no jitter", w=1920, h=1080, spp=64, depth=1, flags=abi.FLAG_NO_JITTER,
               scene=lambda w, h: scenes.random_spheres(1000000, width=w, height=h), rows=[(300, 1), (540, 1), (1000, 1)], windows=[(62, 2)]),
}


def params_for(c, extra_flags=0, **kw):
    kw.setdefault("spp", c["spp"])
    return abi.make_params(c["w"], c["h"], mode=abi.MODE_PATHTRACE, max_depth=c["depth"], seed=1, flags=c["flags"] | extra_flags, **kw)


def bands(c):
    """[(row, first_sample, spp)]: every band row times every sample window."""
    return [(y, s, n) for first, count in c["rows"] for y in range(first, first + count) for s, n in c.get("windows", [(0, c["spp"])])]


def oracle_bands(c, scene, cam):
    """{band: (packed pixels [W], ray segments)} from orc_render_pathtrace, one one-row partition per band."""
    out = {}
    for y, s, n in bands(c):
        frame, _, rays = ol.oracle_pathtrace(scene, cam, params_for(c, tile_rows=1, part_index=y, part_count=c["h"], first_sample=s, spp=n))
        out[(y, s, n)] = (frame[y].copy(), rays)
    return out


def gpu_bands(ctx, c, cam, extra_flags=0):
    """The same bands through the CUDA path, each rendered as the one-row partition of the FULL-SIZE frame."""
    out = {}
    frame = np.zeros((c["h"], c["w"]), np.uint32)
    for y, s, n in bands(c):
        ctx.render(cam, params_for(c, extra_flags, tile_rows=1, part_index=y, part_count=c["h"], first_sample=s, spp=n), out=frame)
        out[(y, s, n)] = (frame[y].copy(), ctx.stats().rays)
    return out


def bands_match(gpu, cpu):
    """Number of differing pixels and whether the per-row ray-segment counts agree."""
    diff = sum(int((gpu[y][0] != cpu[y][0]).sum()) for y in cpu)
    rays_equal = all(gpu[y][1] == cpu[y][1] for y in cpu)
    return diff, rays_equal
