"""Scenes and the comparison used by the independent pin of SURVEY.md section 8 rows a6 / a9 / a10
(analytic spheres, bounce loop + materials, accumulate + resolve), which the reference does not implement.

The checker is oracle/rtiow_book.cpp: the book's program (README.md:2 of the reference names it) in double
precision with its own random numbers; it shares no code with the kernel or with oracle/rt3_oracle.c. The
comparison is therefore statistical. TOLERANCES (also stated in DESIGN.md section 6):

  * converged images, per pixel on the resolved 8-bit frame:       PSNR >= 40 dB (the north star's bar)
  * the same after averaging 4 x 4 pixel blocks (noise / 4):        PSNR >= 46 dB
  * mean linear radiance of the whole image, per channel:           within 0.5 % relative
  * white furnace (albedo 1, uniform white sky): one convex object: every pixel exactly 1; several objects: image
    mean >= 0.99 (paths that reach the depth limit in crevices return 0), no pixel above 1

`render(scene, cam, params) -> (packed frame [H, W] uint32, mean linear radiance [H, W, 3])` is the side under
test: orc_render_pathtrace in tests/test_book_pin.py (CPU), the CUDA path in tests/test_gpu_book_pin.py.
"""
import numpy as np

import oraclelib as ol
from rt3_b200 import abi, scenes

PSNR_PIXEL_DB = 40.0
PSNR_BLOCK_DB = 46.0
MEAN_REL = 5e-3


def materials(rows):
    m = np.zeros(len(rows), abi.MATERIAL_DTYPE)
    for i, (kind, albedo, fuzz, ior) in enumerate(rows):
        m[i]["kind"], m[i]["albedo"], m[i]["fuzz"], m[i]["ior"] = kind, albedo, fuzz, ior
    return m


def sphere_scene(spheres, rows):
    mats = materials(rows)
    return abi.SceneArrays(spheres=np.asarray(spheres, np.float32), sphere_color=np.array([m["albedo"] for m in mats], np.float32),
                           sphere_material=np.arange(len(rows), dtype=np.uint32), materials=mats)


def inside_a_sphere(w, h):
    """The camera sits inside a large Lambertian sphere: every primary ray needs the FAR root of the sphere test; a glass
    ball (far root again, from the inside) and a fuzzy metal ball float in front of it. The shell has a hole-free
    interior, so all light comes from paths that end at the depth limit -> use the uniform sky only through the glass."""
    spheres = [[0, 0, 0, 6.0], [0.8, -0.2, -2.5, 0.7], [-0.9, 0.1, -2.2, 0.6], [0, -1.0, -2.0, 0.3]]
    rows = [(abi.MAT_DIELECTRIC, (1, 1, 1), 0.0, 1.2), (abi.MAT_DIELECTRIC, (1, 1, 1), 0.0, 1.5),
            (abi.MAT_METAL, (0.8, 0.7, 0.4), 0.3, 1.0), (abi.MAT_LAMBERTIAN, (0.7, 0.2, 0.2), 0.0, 1.0)]
    return sphere_scene(spheres, rows), abi.reference_camera(w, h, focal_length=1.0)


def hollow_glass(w, h):
    """C1 plus the book's hollow-glass bubble (negative radius: inward normal)."""
    scene, cam = scenes.rtiow_four_spheres(w, h)
    spheres = np.concatenate([scene.spheres, [[-1, 0, -1, -0.4]]]).astype(np.float32)
    scene = abi.SceneArrays(spheres=spheres, sphere_color=np.concatenate([scene.sphere_color, [[1, 1, 1]]]),
                            sphere_material=np.array([0, 1, 2, 3, 2], np.uint32), materials=scene.materials)
    return scene, cam


def mesh_and_spheres(w, h):
    """Triangles (a tessellated ground quad + a tetrahedron) with all three materials next to analytic spheres."""
    v = np.array([[-4, -0.5, -1], [4, -0.5, -1], [4, -0.5, -9], [-4, -0.5, -9],           # ground quad
                  [-0.6, -0.5, -3.0], [0.6, -0.5, -3.0], [0.0, -0.5, -4.0], [0.0, 0.6, -3.4]], np.float32)  # tetrahedron
    tris = [(0, 1, 2), (0, 2, 3), (4, 5, 7), (5, 6, 7), (6, 4, 7)]
    verts = np.zeros(len(v), abi.VERTEX_DTYPE)
    verts["xyz"] = v
    faces = np.zeros(len(tris), abi.FACE_DTYPE)
    for i, (a, b, c) in enumerate(tris):
        n = np.cross(v[c] - v[a], v[b] - v[a]).astype(np.float64)   # the reference's convention, Triangle.cpp:48
        n /= np.linalg.norm(n)
        faces[i]["v"], faces[i]["normal"], faces[i]["color"] = (a, b, c), n.astype(np.float32), (0.5, 0.5, 0.5)
    mats = materials([(abi.MAT_LAMBERTIAN, (0.6, 0.6, 0.5), 0, 1), (abi.MAT_METAL, (0.9, 0.8, 0.7), 0.1, 1),
                      (abi.MAT_DIELECTRIC, (1, 1, 1), 0, 1.5), (abi.MAT_LAMBERTIAN, (0.2, 0.3, 0.8), 0, 1)])
    scene = abi.SceneArrays(faces=faces, vertices=verts, face_material=np.array([0, 0, 1, 1, 1], np.uint32),
                            spheres=np.array([[1.5, 0.1, -3.5, 0.6], [-1.5, 0.0, -3.0, 0.5]], np.float32),
                            sphere_color=np.ones((2, 3), np.float32), sphere_material=np.array([2, 3], np.uint32), materials=mats)
    return scene, abi.reference_camera(w, h, focal_length=1.5)


CASES = {
    # name: (builder, width, height, spp of both sides, depth, flags)
    "c1_default": (scenes.rtiow_four_spheres, 200, 113, 4096, 50, 0),
    "cover_with_lens": (scenes.rtiow_cover, 120, 80, 512, 50, 0),
    "inside_a_sphere": (inside_a_sphere, 96, 54, 2048, 50, abi.FLAG_UNIFORM_SKY),
    "hollow_glass": (hollow_glass, 128, 72, 2048, 50, 0),
    "mesh_and_spheres": (mesh_and_spheres, 128, 72, 2048, 50, 0),
}


def compare(name, render, spp_scale=1):
    """Renders CASES[name] on the side under test and with the book; asserts the stated tolerances. Returns the figures."""
    build, w, h, spp, depth, flags = CASES[name]
    spp *= spp_scale
    scene, cam = build(w, h)
    params = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=depth, seed=11, flags=flags)
    frame, radiance = render(scene, cam, params)
    book = ol.book_render(scene, cam, w, h, spp, depth, seed=5, flags=flags)
    a, b = ol.unpack_rgb(frame), ol.resolve_8bit(book)
    pix, blk = ol.psnr_levels(a, b), ol.psnr_levels(ol.block_mean(a, 4), ol.block_mean(b, 4))
    rel = np.abs(radiance.astype(np.float64).mean(axis=(0, 1)) / book.astype(np.float64).mean(axis=(0, 1)) - 1).max()
    msg = f"{name}: pixel PSNR {pix:.1f} dB, 4x4-block PSNR {blk:.1f} dB, mean radiance off by {rel:.2e}"
    assert pix >= PSNR_PIXEL_DB and blk >= PSNR_BLOCK_DB and rel <= MEAN_REL, msg
    return pix, blk, rel


FURNACE = {
    # albedo-1 scenes under the uniform white sky: radiance is 1 wherever paths escape before the depth limit
    "lambertian_ball": ([[0, 0, -2, 0.8]], [(abi.MAT_LAMBERTIAN, (1, 1, 1), 0, 1)]),
    "mirror_ball": ([[0, 0, -2, 0.8]], [(abi.MAT_METAL, (1, 1, 1), 0.0, 1)]),
    "glass_ball": ([[0, 0, -2, 0.8]], [(abi.MAT_DIELECTRIC, (1, 1, 1), 0, 1.5)]),
    "default_scene_all_white": ([[0, -100.5, -1, 100], [0, 0, -1, 0.5], [-1, 0, -1, 0.5], [1, 0, -1, 0.5], [-1, 0, -1, -0.4]],
                                [(abi.MAT_LAMBERTIAN, (1, 1, 1), 0, 1), (abi.MAT_LAMBERTIAN, (1, 1, 1), 0, 1), (abi.MAT_DIELECTRIC, (1, 1, 1), 0, 1.5),
                                 (abi.MAT_METAL, (1, 1, 1), 0.0, 1), (abi.MAT_DIELECTRIC, (1, 1, 1), 0, 1.5)]),
}


def furnace(name, render, w=64, h=36, spp=64, depth=50):
    spheres, rows = FURNACE[name]
    scene, cam = sphere_scene(spheres, rows), abi.reference_camera(w, h, focal_length=1.0)
    params = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=depth, seed=2, flags=abi.FLAG_UNIFORM_SKY | abi.FLAG_NO_GAMMA)
    frame, radiance = render(scene, cam, params)
    lo, hi, mean = float(radiance.min()), float(radiance.max()), float(radiance.astype(np.float64).mean())
    assert hi <= 1.0 + 1e-6, f"{name}: radiance up to {hi}: energy was created"
    if name == "default_scene_all_white":
        # several objects: a path caught in a crevice (or in the glass) can reach the depth limit and return 0;
        # that is a few paths in a thousand, the image as a whole stays white within 1 %
        assert mean >= 0.99 and lo >= 0.9, f"{name}: mean radiance {mean}, darkest pixel {lo}"
    else:
        # a single convex object: no path can reach the depth limit, the frame is exactly white
        assert lo == 1.0 and (frame == 0xFFFFFFFF).all(), f"{name}: radiance in [{lo}, {hi}], expected exactly 1"
    return lo, hi
