"""GPU parity, path tracing: CUDA path tracer (through the C ABI) against the CPU restatement.

The reference has no bounce loop / materials / accumulation (PARITY UNPINNED, see
oracle/rt3_oracle.c); the oracle restates RTIOW book-1 semantics. Stated bar
(BASELINE.json): PSNR >= 40 dB at equal spp. Because both sides draw the same
counter-based random numbers and the exact tests and shading use unfused IEEE
arithmetic in the same order, the frames are in fact required to be bit-identical
here, and so are the fixed-point sums behind them (ray counts equal).
"""
import numpy as np
import pytest

import oraclelib as ol
from rt3_b200 import abi, scenes

pytestmark = pytest.mark.gpu


def psnr(a, b):
    ca = np.stack([(a >> s) & 0xFF for s in (24, 16, 8)], -1).astype(np.float64)
    cb = np.stack([(b >> s) & 0xFF for s in (24, 16, 8)], -1).astype(np.float64)
    mse = ((ca - cb) ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def check(ctx, scene, cam, params, exact=True):
    ctx.upload(scene)
    gpu = ctx.render(cam, params)
    st = ctx.stats()
    cpu, _, rays = ol.oracle_pathtrace(scene, cam, params)
    p = psnr(gpu, cpu)
    assert p >= 40.0, f"PSNR {p:.1f} dB < 40 dB"
    if exact:
        assert st.rays == rays, f"ray segments: gpu {st.rays} vs oracle {rays}"
        assert np.array_equal(gpu, cpu), f"{int((gpu != cpu).sum())} pixels differ (PSNR {p:.1f} dB)"
    return gpu


def test_rtiow_four_spheres(gpu_ctx):
    """C1 at reduced size: all three materials, depth 50."""
    w, h = 128, 72
    scene, cam = scenes.rtiow_four_spheres(w, h)
    check(gpu_ctx, scene, cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=16, max_depth=50, seed=1))


def test_cover_scene_with_lens(gpu_ctx):
    """C2's scene (~484 spheres, thin lens) at reduced size."""
    w, h = 96, 64
    scene, cam = scenes.rtiow_cover(w, h)
    check(gpu_ctx, scene, cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=8, max_depth=50, seed=3))


@pytest.mark.parametrize("flags,spp,depth", [(0, 1, 1), (abi.FLAG_NO_JITTER, 1, 1), (abi.FLAG_NO_GAMMA, 4, 3),
                                             (abi.FLAG_NO_JITTER | abi.FLAG_NO_GAMMA, 3, 50)])
def test_flags_and_depths(gpu_ctx, flags, spp, depth):
    w, h = 64, 36
    scene, cam = scenes.rtiow_four_spheres(w, h)
    check(gpu_ctx, scene, cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=depth, seed=9, flags=flags))


def test_depth1_no_jitter_hits_agree_with_reference_mode(gpu_ctx):
    """Deterministic mode: the path tracer's primary visibility is the reference-mode visibility."""
    w, h = 96, 54
    scene, cam = scenes.rtiow_four_spheres(w, h)
    gpu_ctx.upload(scene)
    _, prim, _, _ = gpu_ctx.render_aov(cam, abi.make_params(w, h))
    pt = gpu_ctx.render(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=1, max_depth=1, flags=abi.FLAG_NO_JITTER))
    assert np.array_equal(pt == 0x000000FF, prim != abi.NO_HIT)  # depth 1: hits are black, misses are sky


def test_triangles_and_spheres_with_materials(gpu_ctx):
    from test_gpu_reference_mode import random_soup
    rng = np.random.default_rng(5)
    scene = random_soup(rng, 300, 40)
    mats = np.zeros(6, abi.MATERIAL_DTYPE)
    mats["kind"] = [0, 1, 2, 1, 0, 2]
    mats["albedo"] = rng.uniform(0.2, 0.95, (6, 3))
    mats["fuzz"] = [0, 0.3, 0, 0.0, 0, 0]
    mats["ior"] = [1, 1, 1.5, 1, 1, 1.33]
    scene = abi.SceneArrays(faces=scene.faces, vertices=scene.vertices, face_entity=scene.face_entity,
                            face_material=rng.integers(0, 6, scene.n_faces).astype(np.uint32),
                            spheres=scene.spheres, sphere_color=scene.sphere_color, sphere_entity=scene.sphere_entity,
                            sphere_material=rng.integers(0, 6, scene.n_spheres).astype(np.uint32), materials=mats)
    w, h = 80, 45
    check(gpu_ctx, scene, abi.reference_camera(w, h), abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=6, max_depth=12, seed=2))


def test_streamed_scene(gpu_ctx):
    """More primitives than fit in shared memory: tiles arrive by TMA bulk copy inside the bounce loop."""
    w, h = 48, 27
    scene, cam = scenes.random_spheres(9000, seed=77, width=w, height=h)
    scene.spheres[:, :3] *= np.float32(0.1)   # pull the cloud in so that paths bounce
    scene.spheres[:, 3] *= np.float32(2.0)
    check(gpu_ctx, scene, cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=4, max_depth=6, seed=4))


def test_hollow_glass_sphere(gpu_ctx):
    w, h = 64, 36
    scene, cam = scenes.rtiow_four_spheres(w, h)
    spheres = np.concatenate([scene.spheres, [[-1, 0, -1, -0.4]]]).astype(np.float32)   # negative radius: inward normal
    scene = abi.SceneArrays(spheres=spheres, sphere_color=np.concatenate([scene.sphere_color, [[1, 1, 1]]]),
                            sphere_material=np.array([0, 1, 2, 3, 2], np.uint32), materials=scene.materials)
    check(gpu_ctx, scene, cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=8, max_depth=50, seed=6))


def test_partition_and_schedule_invariance(gpu_ctx):
    """Fixed-point accumulation + per-pixel counters: any row partition gives the identical frame."""
    w, h = 96, 54
    scene, cam = scenes.rtiow_four_spheres(w, h)
    gpu_ctx.upload(scene)
    kw = dict(mode=abi.MODE_PATHTRACE, spp=8, max_depth=20, seed=5)
    full = gpu_ctx.render(cam, abi.make_params(w, h, **kw))
    again = gpu_ctx.render(cam, abi.make_params(w, h, **kw))
    assert np.array_equal(full, again), "render is not run-to-run deterministic"
    for parts, tile in ((2, 8), (8, 4)):
        merged = np.zeros_like(full)
        for i in range(parts):
            gpu_ctx.render(cam, abi.make_params(w, h, tile_rows=tile, part_index=i, part_count=parts, **kw), out=merged)
        assert np.array_equal(merged, full)


def test_full_size_properties(gpu_ctx):
    """BASELINE config C1 at full size (400x225, 100 spp, depth 50): size-independent properties."""
    w, h = 400, 225
    scene, cam = scenes.rtiow_four_spheres(w, h)
    gpu_ctx.upload(scene)
    p = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=100, max_depth=50, seed=1)
    a = gpu_ctx.render(cam, p)
    st = gpu_ctx.stats()
    assert w * h * 100 <= st.rays <= w * h * 100 * 50
    # primary rays are traced against their chunk's candidate list instead of the whole scene: every path has one
    assert 0 < st.beam_rays <= w * h * 100 and 0 < st.beam_tests <= st.beam_rays * 4   # (chunks of path items that span two rows have none)
    assert st.sphere_tests == (st.rays - st.beam_rays) * 4 + st.beam_tests
    assert (a & 0xFF).min() == 0xFF                      # alpha byte
    assert np.array_equal(a, gpu_ctx.render(cam, p))      # idempotent
    top = (a[0] >> 8) & 0xFF                              # sky row: blue channel saturated, gradient intact
    assert top.min() >= 250
    # a strip of the full-size frame equals the oracle on the same strip (row partition = strip)
    strip = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=100, max_depth=50, seed=1, tile_rows=1, part_index=112, part_count=225)
    cpu, _, _ = ol.oracle_pathtrace(scene, cam, strip)
    assert np.array_equal(a[112], cpu[112])


def test_progressive_refinement_equals_one_shot(gpu_ctx):
    """RT3_FLAG_ACCUMULATE (online / progressive mode): k calls of n spp leave exactly the
    frame of one call of k*n spp -- the accumulators are integers -- and every intermediate frame is the one-shot
    frame of the samples so far."""
    w, h = 96, 64
    scene, cam = scenes.rtiow_cover(w, h)
    gpu_ctx.upload(scene)
    base = dict(mode=abi.MODE_PATHTRACE, max_depth=50, seed=5)
    one_shot = {n: gpu_ctx.render(cam, abi.make_params(w, h, spp=n, **base)) for n in (3, 7, 12)}
    rays_12 = gpu_ctx.stats().rays
    rays = 0
    for first, n in ((0, 3), (3, 4), (7, 5)):
        frame = gpu_ctx.render(cam, abi.make_params(w, h, spp=n, first_sample=first, flags=abi.FLAG_ACCUMULATE if first else 0, **base))
        rays += gpu_ctx.stats().rays
        assert np.array_equal(frame, one_shot[first + n]), f"after {first + n} samples"
    assert rays == rays_12
    # a window of samples on its own (no accumulation) is the oracle's frame of those samples
    params = abi.make_params(w, h, spp=4, first_sample=3, **base)
    cpu, _, _ = ol.oracle_pathtrace(scene, cam, params)
    assert np.array_equal(gpu_ctx.render(cam, params), cpu)
    with pytest.raises(abi.Rt3Error, match="ACCUMULATE"):
        gpu_ctx.render(cam, abi.make_params(w + 2, h, spp=1, first_sample=12, flags=abi.FLAG_ACCUMULATE, **base))


def resolve_like_the_kernel(rgb):
    """resolve_kernel on a float image: gamma 2, then glm::packUnorm4x8 = round-half-away(clamp(v, 0, 1) * 255)."""
    v = np.sqrt(rgb.astype(np.float32))
    v = np.clip(v, np.float32(0), np.float32(1)) * np.float32(255)
    ch = np.floor(v.astype(np.float64) + 0.5).astype(np.uint32)
    return (ch[..., 0] << 24) | (ch[..., 1] << 16) | (ch[..., 2] << 8) | np.uint32(0xFF)


def test_radiance_aov_is_the_frame_before_gamma_and_pack(gpu_ctx):
    """rt3_read_radiance: the float image whose resolve is the packed frame, exactly; partitions return their rows only."""
    w, h = 96, 54
    scene, cam = scenes.rtiow_four_spheres(w, h)
    gpu_ctx.upload(scene)
    params = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=6, max_depth=12, seed=5)
    frame = gpu_ctx.render(cam, params)
    rgb = gpu_ctx.read_radiance(w, h)
    assert rgb.shape == (h, w, 3) and np.isfinite(rgb).all() and rgb.min() >= 0 and rgb.max() > 0.5
    assert np.array_equal(resolve_like_the_kernel(rgb), frame)
    with pytest.raises(abi.Rt3Error, match="96x54"):
        gpu_ctx.read_radiance(w, h + 1)
    # the accumulators of a progressive render hold every pass so far
    gpu_ctx.render(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=3, max_depth=12, seed=5))
    gpu_ctx.render(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=3, max_depth=12, seed=5, flags=abi.FLAG_ACCUMULATE, first_sample=3))
    assert np.array_equal(gpu_ctx.read_radiance(w, h), rgb)
    # one partition of three: its rows, zeros elsewhere
    part = gpu_ctx.render(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=6, max_depth=12, seed=5, tile_rows=4, part_index=1, part_count=3))
    mine = (np.arange(h) // 4) % 3 == 1
    got = gpu_ctx.read_radiance(w, h)
    assert np.array_equal(got[mine], rgb[mine]) and not got[~mine].any()
    assert np.array_equal(part[mine], frame[mine])


def test_accumulate_rejects_what_would_blend_different_renders(gpu_ctx):
    """RT3_FLAG_ACCUMULATE only continues the render whose sums the accumulators hold: same scene upload, frame, partition, seed,
    depth, and a sample range that starts where the accumulated one ended (round-1 advice: these were not checked)."""
    w, h = 64, 36
    scene, cam = scenes.rtiow_four_spheres(w, h)
    gpu_ctx.upload(scene)
    base = dict(mode=abi.MODE_PATHTRACE, max_depth=8, seed=5)
    first = abi.make_params(w, h, spp=2, tile_rows=4, part_index=0, part_count=2, **base)

    def cont(**kw):
        args = dict(spp=2, first_sample=2, flags=abi.FLAG_ACCUMULATE, tile_rows=4, part_index=0, part_count=2, **base)
        args.update(kw)
        return abi.make_params(w, h, **args)

    for bad, what in ((cont(part_index=1), "partition"), (cont(tile_rows=2), "partition"), (cont(first_sample=3), "first_sample"),
                      (cont(first_sample=0), "first_sample"), (cont(seed=6), "seed"), (cont(max_depth=9), "seed, max_depth"),
                      (cont(flags=abi.FLAG_ACCUMULATE | abi.FLAG_NO_JITTER), "sampling flags")):
        gpu_ctx.render(cam, first)
        with pytest.raises(abi.Rt3Error, match=what):
            gpu_ctx.render(cam, bad)
    gpu_ctx.render(cam, first)
    gpu_ctx.render(cam, cont())                                   # the right continuation is accepted ...
    gpu_ctx.render(cam, cont(first_sample=4, flags=abi.FLAG_ACCUMULATE | abi.FLAG_BVH))   # ... also through the hierarchy (same sums)
    gpu_ctx.upload(scene)                                          # a new upload invalidates the accumulators
    with pytest.raises(abi.Rt3Error, match="previous path-traced render"):
        gpu_ctx.render(cam, cont(first_sample=6))
    assert gpu_ctx.stats().accel_stack_overflows == 0


def test_primary_candidate_lists_change_nothing(gpu_ctx, monkeypatch):
    """Resident sphere scenes trace their primary rays against a per-chunk candidate list (rt3_kernels.cuh, BEAM kernel): the frame is
    the plain sweep's (RT3_BEAM=0) bit for bit, with thin lens and without, at many and at few samples per pixel (few: a chunk of path
    items spans many pixels or rows, the candidate list grows or is given up), under a row partition, and the ray count is the same."""
    for (w, h, spp, scene_fn, kw) in [(192, 128, 40, scenes.rtiow_cover, {}), (192, 128, 3, scenes.rtiow_cover, {}), (160, 90, 64, scenes.rtiow_four_spheres, {}),
                                      (97, 61, 7, scenes.rtiow_cover, dict(tile_rows=3, part_index=1, part_count=2)), (64, 40, 1, scenes.rtiow_four_spheres, {})]:
        scene, cam = scene_fn(w, h)
        gpu_ctx.upload(scene)
        p = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=50, seed=3, **kw)
        monkeypatch.setenv("RT3_BEAM", "0")
        plain = gpu_ctx.render(cam, p).copy()
        st0 = gpu_ctx.stats()
        assert st0.beam_rays == 0 and st0.sphere_tests == st0.rays * scene.n_spheres
        monkeypatch.delenv("RT3_BEAM")
        beam = gpu_ctx.render(cam, p)
        st1 = gpu_ctx.stats()
        assert np.array_equal(plain, beam)
        assert st1.rays == st0.rays and st1.beam_rays <= st1.rays
        if spp >= 40:
            assert st1.beam_rays > 0 and st1.beam_tests < st1.beam_rays * scene.n_spheres
