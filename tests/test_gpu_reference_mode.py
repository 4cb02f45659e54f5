"""GPU parity, reference mode: the CUDA ray caster (through the C ABI) against
the reference's golden vectors and the pinned CPU restatement.

Bar (BASELINE.json north_star, deterministic mode): closest-hit ids bit-exact,
hit distances within 1e-5 relative -- the kernels are built without FMA
contraction in the exact tests, so t is in fact required to be bit-exact here;
packed pixels bit-exact.
"""
import numpy as np
import pytest

import oraclelib as ol
from conftest import load_golden
from rt3_b200 import abi, scenes

pytestmark = pytest.mark.gpu


def check_against_oracle(ctx, scene, cam, w, h):
    ctx.upload(scene)
    params = abi.make_params(w, h, mode=abi.MODE_REFERENCE)
    frame, prim, ent, t = ctx.render_aov(cam, params)
    oframe, oprim, oent, ot = ol.oracle_reference(scene, cam, w, h)
    assert np.array_equal(prim, oprim), f"{int((prim != oprim).sum())} primitive ids differ"
    assert np.array_equal(ent, oent)
    hit = oprim != abi.NO_HIT
    assert np.all(np.isinf(t[~hit])) and np.all(t[~hit] > 0)
    rel = np.abs(t[hit].astype(np.float64) - ot[hit]) / np.maximum(np.abs(ot[hit]), 1e-30)
    assert rel.size == 0 or rel.max() <= 1e-5
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32)), "t is expected bit-exact (no FMA in the exact tests)"
    assert np.array_equal(frame, oframe)
    assert np.array_equal(ctx.render(cam, params), oframe)
    return frame


@pytest.mark.parametrize("name", ["triangle_400x225", "sphere8_400x225", "default_400x225"])
def test_golden_scenes(gpu_ctx, name):
    g, scene = load_golden(name)
    h, w = g["frame"].shape[0] + 1, g["frame"].shape[1]
    gpu_ctx.upload(scene)
    frame, prim, ent, t = gpu_ctx.render_aov(abi.reference_camera(w, h), abi.make_params(w, h))
    assert np.array_equal(frame[:h - 1], g["frame"])
    assert np.array_equal(prim[:h - 1], g["prim"])
    assert np.array_equal(t[:h - 1].view(np.uint32), g["t_bits"])
    hit = prim != abi.NO_HIT
    assert np.array_equal(ent[hit], scene.face_entity[prim[hit]])


def random_soup(rng, n_faces, n_spheres, degenerate=False):
    verts = np.zeros(3 * n_faces, abi.VERTEX_DTYPE)
    centres = rng.uniform([-3, -2, -9], [3, 2, -2], (n_faces, 1, 3))
    verts["xyz"] = (centres + rng.normal(0, 0.5, (n_faces, 3, 3))).reshape(-1, 3).astype(np.float32)
    faces = np.zeros(n_faces, abi.FACE_DTYPE)
    faces["v"] = np.arange(3 * n_faces, dtype=np.uint32).reshape(-1, 3)
    p = verts["xyz"].reshape(-1, 3, 3)
    # normal as the reference computes it: normalize(cross(p3 - p1, p2 - p1)) in fp32 (Triangle.cpp:48)
    a, b = p[:, 2] - p[:, 0], p[:, 1] - p[:, 0]
    c = np.stack([a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2], a[:, 2] * b[:, 0] - b[:, 2] * a[:, 0], a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]], 1).astype(np.float32)
    inv = (np.float32(1.0) / np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2], dtype=np.float32)).astype(np.float32)
    faces["normal"] = c * inv[:, None]
    faces["color"] = rng.uniform(0, 1, (n_faces, 3)).astype(np.float32)
    if degenerate and n_faces >= 4:
        verts["xyz"][3:6] = verts["xyz"][3]         # zero-area face -> NaN normal: must never be hit
        faces["normal"][1] = np.nan
        faces["normal"][2] = 0.0                    # n.d == 0 for every ray
        faces["v"][3] = faces["v"][0]               # exact duplicate of face 0: ties keep the lower index
        faces["normal"][3] = faces["normal"][0]
    spheres = np.concatenate([rng.uniform([-3, -2, -9], [3, 2, -2], (n_spheres, 3)), rng.uniform(0.1, 0.8, (n_spheres, 1))], 1).astype(np.float32)
    return abi.SceneArrays(faces=faces, vertices=verts, face_entity=rng.integers(0, 9, n_faces).astype(np.uint32),
                           spheres=spheres, sphere_color=rng.uniform(0, 1, (n_spheres, 3)).astype(np.float32),
                           sphere_entity=rng.integers(9, 20, n_spheres).astype(np.uint32))


@pytest.mark.parametrize("n_faces,n_spheres,w,h,degenerate", [
    (0, 0, 64, 36, False),        # empty scene: sky only
    (1, 0, 33, 17, False),
    (0, 1, 33, 17, False),
    (40, 0, 97, 61, True),        # NaN / zero normals, duplicate faces
    (31, 1, 64, 36, False),       # exactly one prefilter block
    (33, 31, 64, 36, False),      # ragged second block
    (700, 300, 160, 90, False),
    (700, 68, 128, 72, False),    # largest constant-bank scene (768 primitives)
    (700, 69, 128, 72, False),    # smallest streamed scene
    (4000, 96, 128, 72, False),   # exactly four full streamed tiles
    (5000, 1500, 128, 72, False), # streamed tiles (TMA bulk copies), several tiles, ragged tail
])
def test_random_scenes_match_oracle(gpu_ctx, n_faces, n_spheres, w, h, degenerate):
    rng = np.random.default_rng(n_faces * 7919 + n_spheres)
    check_against_oracle(gpu_ctx, random_soup(rng, n_faces, n_spheres, degenerate), abi.reference_camera(w, h), w, h)


def test_analytic_sphere_cloud_depth1(gpu_ctx):
    """C5-shaped case at oracle-sized scale: random sphere cloud, un-jittered primary rays."""
    w, h = 160, 90
    scene, cam = scenes.random_spheres(20000, width=w, height=h)
    check_against_oracle(gpu_ctx, scene, cam, w, h)


def test_partitions_reassemble_to_the_single_device_frame(gpu_ctx):
    g, scene = load_golden("sphere8_400x225")
    w, h = 400, 225
    gpu_ctx.upload(scene)
    cam = abi.reference_camera(w, h)
    full = gpu_ctx.render(cam, abi.make_params(w, h))
    for parts, tile in ((2, 8), (3, 5), (8, 16)):
        merged = np.zeros_like(full)
        rows = 0
        for i in range(parts):
            merged_before = merged.copy()
            gpu_ctx.render(cam, abi.make_params(w, h, tile_rows=tile, part_index=i, part_count=parts), out=merged)
            rows += gpu_ctx.stats().rows_rendered
            own = ((np.arange(h) // tile) % parts) == i
            assert np.array_equal(merged[~own], merged_before[~own]), "a partition wrote rows it does not own"
        assert rows == h and np.array_equal(merged, full)


def test_errors_are_reported_not_swallowed(gpu_ctx):
    ctx = abi.Context(0)
    with pytest.raises(abi.Rt3Error, match="rt3_scene_upload"):
        ctx.render(abi.reference_camera(8, 8), abi.make_params(8, 8))
    g, scene = load_golden("triangle_400x225")
    bad = abi.SceneArrays(faces=scene.faces.copy(), vertices=scene.vertices[:2].copy())
    with pytest.raises(abi.Rt3Error, match="out of range"):
        ctx.upload(bad)
    ctx.upload(scene)
    with pytest.raises(abi.Rt3Error, match="2x2"):
        ctx.render(abi.reference_camera(8, 8), abi.make_params(1, 8))
    with pytest.raises(abi.Rt3Error, match="spp"):
        ctx.render(abi.reference_camera(8, 8), abi.make_params(8, 8, mode=abi.MODE_PATHTRACE, spp=0))
    ctx.close()


def test_two_contexts_share_the_constant_bank_correctly(gpu_ctx):
    """Scenes up to RT3_CONST_PRIMS are swept out of one constant-memory bank per device; contexts with different
    scenes rendering in turn must each get their own records back (ownership check in claim_constant_bank)."""
    w, h = 64, 36
    cam = abi.reference_camera(w, h)
    other = abi.Context(0)
    try:
        scenes_ab = [random_soup(np.random.default_rng(s), 40, 60) for s in (21, 22)]
        want = [ol.oracle_reference(sc, cam, w, h)[0] for sc in scenes_ab]
        gpu_ctx.upload(scenes_ab[0])
        other.upload(scenes_ab[1])
        for _ in range(3):
            assert np.array_equal(gpu_ctx.render(cam, abi.make_params(w, h)), want[0])
            assert np.array_equal(other.render(cam, abi.make_params(w, h)), want[1])
            pt = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=2, max_depth=4, seed=1)
            a, b = gpu_ctx.render(cam, pt), other.render(cam, pt)
            assert np.array_equal(a, gpu_ctx.render(cam, pt)) and np.array_equal(b, other.render(cam, pt)) and not np.array_equal(a, b)
    finally:
        other.close()
