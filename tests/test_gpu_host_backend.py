"""GPU: the C++ host backend end to end (ECS::create_* -> CudaRenderer::prerender -> render -> Frame),
and the drop-in proof: the reference's UNMODIFIED main() linked against CudaRenderer renders the
reference's default scene to the reference's golden image."""
import os
import subprocess

import numpy as np
import pytest

import hostlib
import oraclelib as ol
from conftest import ROOT, load_golden
from rt3_b200 import abi, scenes

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def test_reference_mode_through_the_host_backend(built):
    g, _ = load_golden("sphere8_400x225")
    hs = hostlib.HostScene()
    hs.add_sphere((0, 0, -3), 1.0, 8, 8, (1, 0, 0))
    hs.create_renderer(mode=abi.MODE_REFERENCE)
    hs.prerender()
    frame, ms, rays = hs.render(400, 225)
    assert np.array_equal(frame[:224], g["frame"])
    assert rays == 400 * 225 and ms > 0
    # render() may be called again with another camera after one prerender() (Renderer.hpp:48-50)
    frame2, _, _ = hs.render(200, 113)
    small, _, _, _ = ol.oracle_reference(hs.flatten(), abi.reference_camera(200, 113), 200, 113)
    assert np.array_equal(frame2, small)


def test_path_tracing_through_the_host_backend(built):
    """C1 built through ECS::create_sphere + the material side table equals the oracle on the same scene."""
    w, h = 96, 54
    hs = hostlib.HostScene()
    for c, r, col in (((0, -100.5, -1), 100.0, (0.8, 0.8, 0.0)), ((0, 0, -1), 0.5, (0.1, 0.2, 0.5)),
                      ((-1, 0, -1), 0.5, (1, 1, 1)), ((1, 0, -1), 0.5, (0.8, 0.6, 0.2))):
        hs.add_sphere(c, r, 8, 8, col)
    hs.create_renderer(mode=abi.MODE_PATHTRACE, spp=8, max_depth=50, seed=3, analytic_spheres=True)
    hs.set_material(2, abi.MAT_DIELECTRIC, (1, 1, 1), 0.0, 1.5)
    hs.set_material(3, abi.MAT_METAL, (0.8, 0.6, 0.2), 0.0, 1.0)
    hs.prerender()
    frame, _, rays = hs.render(w, h, focal=1.0)
    scene, cam = scenes.rtiow_four_spheres(w, h)
    cpu, _, cpu_rays = ol.oracle_pathtrace(scene, cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=8, max_depth=50, seed=3))
    assert rays == cpu_rays and np.array_equal(frame, cpu)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "raytracer_cuda")), reason="drop-in binary not built (needs /root/reference at build time)")
def test_reference_main_with_cuda_backend_renders_the_golden_image(tmp_path):
    g, _ = load_golden("default_400x225")
    out = tmp_path / "out.ppm"
    subprocess.run([os.path.join(REF_DIR, "raytracer_cuda"), "-W", "400", "-H", "225", "-f", "ppm", str(out)], cwd=REF_DIR, check=True,
                   stdout=subprocess.DEVNULL, timeout=120)
    from PIL import Image   # the reference's writer puts a comment line in the header (Frame.cpp:127)
    rgb = np.array(Image.open(out).convert("RGB")).astype(np.uint32)
    assert rgb.shape == (225, 400, 3)
    frame = (rgb[..., 0] << 24) | (rgb[..., 1] << 16) | (rgb[..., 2] << 8) | 0xFF
    assert np.array_equal(frame[:224], g["frame"])


@pytest.mark.parametrize("channels", [3, 4])
def test_device_frame_to_bytes(gpu_ctx, channels):
    """rt3_frame_bytes (output side, reference Frame.cpp:88-96,131-142) against the same unpacking in numpy."""
    import torch
    rng = np.random.default_rng(channels)
    for w, h in ((1, 1), (5, 3), (64, 36), (401, 227)):   # pixel counts with every remainder modulo 4
        frame = rng.integers(0, 1 << 32, (h, w), dtype=np.uint32)
        d_frame = torch.from_numpy(frame.view(np.int32)).cuda()
        d_out = torch.zeros(w * h * channels + 16, dtype=torch.uint8, device="cuda")
        gpu_ctx.frame_bytes(d_frame.data_ptr(), d_out.data_ptr(), w, h, channels)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        want = np.stack([(frame >> s) & 0xFF for s in (24, 16, 8)] + ([np.full_like(frame, 255)] if channels == 4 else []), -1).astype(np.uint8)
        assert np.array_equal(got[: w * h * channels], want.ravel()) and not got[w * h * channels:].any()
    with pytest.raises(abi.Rt3Error, match="channels"):
        gpu_ctx.frame_bytes(d_frame.data_ptr(), d_out.data_ptr(), w, h, 2)


def test_progressive_mode_through_the_host_backend(built):
    """CudaRenderer::render_progressive: pass k's frame is the one-shot frame of (k + 1) * spp samples."""
    w, h = 80, 45
    def scene(spp):
        hs = hostlib.HostScene()
        n, _ = hs.add_scene_text('''entities {
            sphere ground { center: 0.0 -100.5 -1.0; radius: 100.0; n_meridians: 8; n_parallels: 8; color: 0.8 0.8 0.0; }
            sphere ball   { center: 0.0 0.0 -1.0;    radius: 0.5;   n_meridians: 8; n_parallels: 8; color: 0.1 0.2 0.5; } }''')
        assert n == 2
        hs.create_renderer(mode=abi.MODE_PATHTRACE, spp=spp, max_depth=20, seed=4, analytic_spheres=True)
        hs.prerender()
        return hs
    frames = scene(3).render_progressive(w, h, 4, focal=1.0)
    for k in (0, 1, 3):
        one_shot, _, _ = scene(3 * (k + 1)).render(w, h, focal=1.0)
        assert np.array_equal(frames[k], one_shot), f"pass {k}"
    assert not np.array_equal(frames[0], frames[3])


def gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", [abi.MODE_REFERENCE, abi.MODE_PATHTRACE])
def test_several_devices_render_the_one_device_frame(built, mode):
    """CudaRenderer over several contexts (SURVEY 8e behind the reference's Renderer interface): every context renders its
    row tiles into the first one's frame. Uses distinct GPUs when the box has them; otherwise the same GPU three times,
    which runs the same split, frame sharing and statistics merge."""
    w, h = 120, 67
    devices = [0, 1, 0] if gpu_count() >= 2 else [0, 0, 0]

    def scene(multi):
        hs = hostlib.HostScene()
        hs.add_sphere((0, -100.5, -1), 100.0, 8, 8, (0.8, 0.8, 0.0))
        hs.add_sphere((0, 0, -1), 0.5, 12, 12, (0.1, 0.2, 0.5))
        hs.add_triangle((1, 0, -2), (-1, 0, -2), (0, 1, -2), (1, 0, 0))
        kw = dict(mode=mode, spp=5, max_depth=12, seed=9)
        if multi:
            hs.create_renderer_multi(devices, tile_rows=4, **kw)
        else:
            hs.create_renderer(**kw)
        hs.prerender()
        return hs

    one, _, rays_one = scene(False).render(w, h, focal=1.0)
    hs = scene(True)
    many, ms, rays_many = hs.render(w, h, focal=1.0)
    assert np.array_equal(many, one)
    assert rays_many == rays_one and ms > 0
    again, _, _ = hs.render(w // 2, h // 2, focal=1.0)        # the shared frame is re-allocated for another size
    small, _, _ = scene(False).render(w // 2, h // 2, focal=1.0)
    assert np.array_equal(again, small)
    if mode == abi.MODE_PATHTRACE:
        frames = hs.render_progressive(w, h, 2, focal=1.0)     # accumulators stay with the context that owns the rows
        assert np.array_equal(frames[0], one)
        hs2 = hostlib.HostScene()
        hs2.add_sphere((0, -100.5, -1), 100.0, 8, 8, (0.8, 0.8, 0.0))
        hs2.add_sphere((0, 0, -1), 0.5, 12, 12, (0.1, 0.2, 0.5))
        hs2.add_triangle((1, 0, -2), (-1, 0, -2), (0, 1, -2), (1, 0, 0))
        hs2.create_renderer(mode=mode, spp=10, max_depth=12, seed=9)
        hs2.prerender()
        ten, _, _ = hs2.render(w, h, focal=1.0)
        assert np.array_equal(frames[1], ten)


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline() == b"PF\n"
        w, h = map(int, f.readline().split())
        assert float(f.readline()) < 0          # little-endian
        data = np.frombuffer(f.read(), "<f4").reshape(h, w, 3)
    return data[::-1]                            # rows are stored bottom to top


def test_radiance_dump_through_the_host_backend(built, tmp_path):
    """CudaRenderer::read_radiance / write_radiance_pfm (float AOV of the output side), one and several contexts."""
    w, h = 64, 36

    def scene(devices):
        hs = hostlib.HostScene()
        hs.add_sphere((0, -100.5, -1), 100.0, 8, 8, (0.8, 0.8, 0.0))
        hs.add_sphere((0, 0, -1), 0.5, 8, 8, (0.1, 0.2, 0.5))
        kw = dict(mode=abi.MODE_PATHTRACE, spp=4, max_depth=10, seed=2, analytic_spheres=True)
        if devices:
            hs.create_renderer_multi(devices, tile_rows=2, **kw)
        else:
            hs.create_renderer(**kw)
        hs.prerender()
        return hs

    hs = scene(None)
    frame, _, _ = hs.render(w, h, focal=1.0)
    path = str(tmp_path / "radiance.pfm")
    rgb = hs.radiance(w, h, path)
    v = np.clip(np.sqrt(rgb), np.float32(0), np.float32(1)) * np.float32(255)
    ch = np.floor(v.astype(np.float64) + 0.5).astype(np.uint32)
    assert np.array_equal((ch[..., 0] << 24) | (ch[..., 1] << 16) | (ch[..., 2] << 8) | np.uint32(0xFF), frame)
    assert np.array_equal(read_pfm(path), rgb)
    many = scene([0, 0])
    frame2, _, _ = many.render(w, h, focal=1.0)
    assert np.array_equal(frame2, frame) and np.array_equal(many.radiance(w, h), rgb)
    ref = hostlib.HostScene()
    ref.add_triangle((1, 0, -3), (-1, 0, -3), (0, 1, -3), (1, 0, 0))
    ref.create_renderer(mode=abi.MODE_REFERENCE)
    ref.prerender()
    ref.render(w, h)
    with pytest.raises(hostlib.HostError, match="path-traced"):
        ref.radiance(w, h)


def test_cli_over_several_contexts_and_pfm(built, tmp_path):
    """rt3_render (initialize_renderer + RT3_DEVICES): the frame from two contexts equals the one-context frame; --pfm writes the float image."""
    exe = os.path.join(ROOT, "raytracer-3_b200", "host", "rt3_render")
    outs = {}
    for tag, devices in (("one", None), ("two", "0,0")):
        env = dict(os.environ)
        env.pop("RT3_DEVICES", None)
        if devices:
            env["RT3_DEVICES"] = devices
        ppm, pfm = tmp_path / f"{tag}.ppm", tmp_path / f"{tag}.pfm"
        subprocess.run([exe, "-s", "rtiow", "-W", "160", "-H", "90", "--spp", "4", "--depth", "10", "-o", str(ppm), "--pfm", str(pfm)],
                       check=True, env=env, stdout=subprocess.DEVNULL, timeout=120)
        outs[tag] = (ppm.read_bytes(), read_pfm(str(pfm)))
    assert outs["one"][0] == outs["two"][0]
    assert outs["one"][1].shape == (90, 160, 3) and np.array_equal(outs["one"][1], outs["two"][1])
