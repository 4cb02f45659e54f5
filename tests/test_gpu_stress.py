"""GPU: the conservative prefilter and the hierarchy under hostile scales.

Both stand in front of the exact tests and may only ever add candidates. Their slack terms
are derived from rounding bounds in |c|, |o| and r; these scenes push those ratios far beyond the BASELINE configs:
coordinates of 10^3..10^5 with radii down to 10^-2, cameras far from the origin, grazing rays along a row of spheres,
sliver triangles, everything shifted so that the scene basis is oblique. Required: ids, t bits and pixels identical to
the CPU oracle for the sweep and for the hierarchy.
"""
import numpy as np
import pytest

import oraclelib as ol
from rt3_b200 import abi
from test_gpu_reference_mode import random_soup

pytestmark = pytest.mark.gpu


def look_camera(origin, target, w, h, fov_scale=1.0):
    o, t = np.array(origin, np.float64), np.array(target, np.float64)
    f = (t - o) / np.linalg.norm(t - o)
    up = np.array([0.0, 1.0, 0.0]) if abs(f[1]) < 0.9 else np.array([1.0, 0.0, 0.0])
    r = np.cross(f, up); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    hor, ver = 2.0 * fov_scale * (w / h) * r, 2.0 * fov_scale * u
    llc = o - hor / 2 - ver / 2 + 2.0 * f
    return abi.make_camera(o, hor, ver, llc)


def both_paths_match_oracle(ctx, scene, cam, w, h):
    ctx.upload(scene)
    oframe, oprim, oent, ot = ol.oracle_reference(scene, cam, w, h)
    for flags in (0, abi.FLAG_BVH):
        frame, prim, ent, t = ctx.render_aov(cam, abi.make_params(w, h, flags=flags))
        bad = prim != oprim
        assert not bad.any(), f"flags={flags}: {int(bad.sum())} ids differ, e.g. pixel {np.argwhere(bad)[0]} gpu {prim[bad][0]} oracle {oprim[bad][0]}"
        assert np.array_equal(t.view(np.uint32), ot.view(np.uint32)) and np.array_equal(frame, oframe)
    return oprim


@pytest.mark.parametrize("offset,spread,rmin,rmax", [
    ((0, 0, -30), 10, 0.01, 0.05),            # tiny spheres
    ((1000, -2000, 3000), 40, 0.05, 2.0),     # far from the origin
    ((3e4, 1e4, -2e4), 300, 0.5, 30.0),       # very far: float spacing of the coordinates ~ 2e-3
    ((0, 0, 0), 5, 0.5, 200.0),               # huge spheres around the camera (rays start inside some)
])
def test_sphere_clouds_at_hostile_scales(gpu_ctx, offset, spread, rmin, rmax):
    rng = np.random.default_rng(int(abs(offset[0]) + spread))
    n, w, h = 600, 96, 54
    c = np.array(offset) + rng.normal(0, spread, (n, 3))
    spheres = np.concatenate([c, rng.uniform(rmin, rmax, (n, 1))], 1).astype(np.float32)
    scene = abi.SceneArrays(spheres=spheres, sphere_color=rng.uniform(0, 1, (n, 3)).astype(np.float32), sphere_entity=np.arange(n, dtype=np.uint32))
    eye = np.array(offset) + np.array([0.3 * spread, 0.2 * spread, 3.0 * spread])
    hit = both_paths_match_oracle(gpu_ctx, scene, look_camera(eye, offset, w, h, 0.6), w, h)
    assert (hit != abi.NO_HIT).mean() > 0.002


def test_grazing_rays_along_a_row(gpu_ctx):
    """A thousand unit spheres on a line, seen almost along the line: every ray passes within a hair of many of them."""
    n, w, h = 1000, 128, 32
    x = np.arange(n, dtype=np.float64) * 2.5
    spheres = np.stack([x, np.full(n, 1e-3), np.full(n, -5.0), np.ones(n)], 1).astype(np.float32)
    scene = abi.SceneArrays(spheres=spheres, sphere_entity=np.arange(n, dtype=np.uint32))
    hit = both_paths_match_oracle(gpu_ctx, scene, look_camera((-30.0, 4.0, -5.0), (2500.0, 1.0, -5.0), w, h, 0.03), w, h)
    assert len(np.unique(hit)) > 8


def test_slivers_and_shifted_mesh(gpu_ctx):
    rng = np.random.default_rng(11)
    scene = random_soup(rng, 3000, 200)
    # stretch every triangle into a sliver along a random direction and move the whole scene far away along an oblique axis
    v = scene.vertices["xyz"].reshape(-1, 3, 3).astype(np.float64)
    centre = v.mean(1, keepdims=True)
    axis = rng.normal(0, 1, (len(v), 1, 3)); axis /= np.linalg.norm(axis, axis=2, keepdims=True)
    d = v - centre
    v = centre + d * 0.02 + axis * (d * axis).sum(2, keepdims=True) * 40.0
    shift = np.array([700.0, -350.0, 1200.0])
    scene.vertices["xyz"] = (v + shift).reshape(-1, 3).astype(np.float32)
    p = scene.vertices["xyz"].reshape(-1, 3, 3)
    a, b = p[:, 2] - p[:, 0], p[:, 1] - p[:, 0]
    c = np.stack([a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2], a[:, 2] * b[:, 0] - b[:, 2] * a[:, 0], a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]], 1).astype(np.float32)
    inv = (np.float32(1.0) / np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2], dtype=np.float32)).astype(np.float32)
    scene.faces["normal"] = c * inv[:, None]
    scene.spheres[:, :3] += shift.astype(np.float32)
    w, h = 120, 68
    hit = both_paths_match_oracle(gpu_ctx, scene, look_camera(shift + np.array([0.0, 0.0, 2.0]), shift + np.array([0.0, 0.0, -6.0]), w, h), w, h)
    assert (hit != abi.NO_HIT).mean() > 0.05


def test_path_tracing_far_from_the_origin(gpu_ctx):
    """Bounce rays start on surfaces, so |o| is large for every ray after the first: the per-ray slack must hold."""
    rng = np.random.default_rng(2)
    n, w, h = 300, 64, 36
    offset = np.array([5000.0, 800.0, -2500.0])
    c = offset + rng.normal(0, 6, (n, 3))
    spheres = np.concatenate([c, rng.uniform(0.2, 1.5, (n, 1))], 1).astype(np.float32)
    mats = np.zeros(3, abi.MATERIAL_DTYPE)
    mats["kind"] = [abi.MAT_LAMBERTIAN, abi.MAT_METAL, abi.MAT_DIELECTRIC]
    mats["albedo"] = [(0.7, 0.7, 0.7), (0.9, 0.8, 0.6), (1, 1, 1)]
    mats["fuzz"] = [0, 0.05, 0]; mats["ior"] = [1, 1, 1.5]
    scene = abi.SceneArrays(spheres=spheres, sphere_material=rng.integers(0, 3, n).astype(np.uint32), sphere_entity=np.arange(n, dtype=np.uint32), materials=mats)
    cam = look_camera(offset + np.array([2.0, 1.0, 25.0]), offset, w, h, 0.5)
    gpu_ctx.upload(scene)
    base = dict(mode=abi.MODE_PATHTRACE, spp=8, max_depth=30, seed=6)
    cpu, _, rays = ol.oracle_pathtrace(scene, cam, abi.make_params(w, h, **base))
    for flags in (0, abi.FLAG_BVH):
        gpu = gpu_ctx.render(cam, abi.make_params(w, h, flags=flags, **base))
        assert gpu_ctx.stats().rays == rays and np.array_equal(gpu, cpu), f"flags={flags}: {int((gpu != cpu).sum())} pixels differ"


def test_two_host_threads_share_the_constant_bank(built):
    """Two contexts on one device, each driven by its own host thread with its own small scene (both small enough for
    the constant-bank sweep): the bank is claimed under a per-device lock that is held until the kernel reading it has been
    enqueued, and a change of owner waits for the device, so neither thread can sweep the other's records."""
    import threading
    from rt3_b200 import scenes
    w, h = 160, 90
    jobs = []
    for seed in (1, 2):
        scene, cam = scenes.random_spheres(300, seed=seed, width=w, height=h)
        scene.spheres[:, :3] *= np.float32(0.05)
        scene.spheres[:, 2] -= np.float32(1.0)
        ctx = abi.Context(0)
        ctx.upload(scene)
        params = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=4, max_depth=6, seed=seed)
        jobs.append((ctx, cam, params, ctx.render(cam, params)))   # the frame each context renders on its own
    assert not np.array_equal(jobs[0][3], jobs[1][3])
    errors = []

    def hammer(ctx, cam, params, expected):
        try:
            for _ in range(40):
                if not np.array_equal(ctx.render(cam, params), expected):
                    errors.append("a frame differs from the one rendered alone")
                    return
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=hammer, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for ctx, *_ in jobs:
        ctx.close()
    assert not errors, errors
