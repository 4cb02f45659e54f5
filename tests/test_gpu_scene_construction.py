"""GPU: scene construction on the device against the reference's CPU pre-render.

rt3_tessellate_spheres restates cpu_pre_render_sphere (reference src/lib/entities/Sphere.cpp:69-79,120-351) as
CUDA kernels. The host backend's tessellation is bit-equal to the compiled reference's
(tests/test_host_backend.py), so it is the checker here. Indices must be identical; coordinates come from
double-precision sin / cos narrowed to float, where the device's libm may differ from glibc's in the last
place of the double -- the bar is 1 float ulp per coordinate (the reference's own GPU shader is ~1e-6 away from
its CPU path), and the cases below are in fact required to be bit-identical.
"""
import numpy as np
import pytest

import hostlib
from conftest import load_golden
from rt3_b200 import abi

pytestmark = pytest.mark.gpu

SPHERES = [((0, 0, -3), 1.0, 8, 8, (1, 0, 0), 0), ((0.6, 0.2, -4), 0.9, 12, 9, (0.2, 0.9, 0.3), 1), ((0.1, 0.2, -2), 0.3, 5, 3, (0.2, 0.9, 0.3), 2),
           ((-2, 0, -5), 1.0, 1, 3, (0, 0, 1), 3), ((3, 1, -7), 2.5, 64, 33, (0.9, 0.9, 0.1), 4)]


def ulp_distance(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.where(np.isnan(a) & np.isnan(b), 0, np.abs(ia - ib))   # degenerate faces (m = 1): NaN normals on both sides


def host_flatten(spheres):
    hs = hostlib.HostScene()
    for center, radius, m, p, color, _ in spheres:
        hs.add_sphere(center, radius, m, p, color)
    return hs.flatten()


def test_batch_matches_the_cpu_pre_render(gpu_ctx):
    dev = gpu_ctx.tessellate_spheres(SPHERES)
    cpu = host_flatten(SPHERES)
    assert dev.n_faces == cpu.n_faces == sum(2 * m * (p - 2) for _, _, m, p, _, _ in SPHERES)
    assert np.array_equal(dev.faces["v"], cpu.faces["v"]), "face indices (absolute, entity after entity) differ"
    assert np.array_equal(dev.face_entity, cpu.face_entity)
    for name, a, b in (("vertices", dev.vertices["xyz"], cpu.vertices["xyz"]), ("normals", dev.faces["normal"], cpu.faces["normal"]),
                       ("colours", dev.faces["color"], cpu.faces["color"])):
        d = ulp_distance(np.ascontiguousarray(a), np.ascontiguousarray(b))
        assert d.max() <= 1, f"{name}: {d.max()} ulp"
        assert d.max() == 0, f"{name}: {int((d > 0).sum())} values differ in the last place"


def test_c3_mesh_and_index_offset(gpu_ctx):
    """create_sphere(.., 225, 225, ..): 100 350 faces; indices shifted like transfer_entity (SequentialRenderer.cpp:181-187)."""
    one = [((0, 0, -3), 1.0, 225, 225, (1, 0, 0), 7)]
    dev = gpu_ctx.tessellate_spheres(one, first_vertex=1000)
    cpu = host_flatten(one)
    assert (dev.n_faces, len(dev.vertices)) == (100350, 50177)
    assert np.array_equal(dev.faces["v"], cpu.faces["v"] + 1000) and np.all(dev.face_entity == 7)
    d = ulp_distance(np.ascontiguousarray(dev.vertices["xyz"]), np.ascontiguousarray(cpu.vertices["xyz"]))
    assert d.max() <= 1
    n = ulp_distance(np.ascontiguousarray(dev.faces["normal"]), np.ascontiguousarray(cpu.faces["normal"]))
    assert d.max() == 0 and n.max() == 0, f"{int((d > 0).sum())} coordinates / {int((n > 0).sum())} normal components differ"


def test_host_backend_with_device_tessellation_renders_the_golden_image(built):
    g, _ = load_golden("sphere8_400x225")
    hs = hostlib.HostScene()
    hs.add_sphere((0, 0, -3), 1.0, 8, 8, (1, 0, 0))
    hs.create_renderer(mode=abi.MODE_REFERENCE, device_tessellation=True)
    hs.prerender()
    frame, _, _ = hs.render(400, 225)
    assert np.array_equal(frame[:224], g["frame"])


def test_bad_arguments(gpu_ctx):
    with pytest.raises(abi.Rt3Error, match="n_parallels"):
        gpu_ctx.tessellate_spheres([((0, 0, 0), 1.0, 8, 2, (1, 1, 1), 0)])
    assert gpu_ctx.tessellate_spheres([]).n_faces == 0
