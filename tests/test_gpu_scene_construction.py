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


# ---- the derived arrays are built on the device: rt3_scene_upload (host arrays) and rt3_scene_upload_device (device arrays) ----

class DeviceScene:
    """A SceneArrays copied into rt3_buffer_alloc'ed device arrays."""

    def __init__(self, ctx, scene):
        self.ctx, self.scene = ctx, scene
        names = ("faces", "vertices", "face_material", "face_entity", "spheres", "sphere_color", "sphere_material", "sphere_entity", "materials")
        self.ptr = {n: ctx.to_device(getattr(scene, n)) for n in names}

    def upload(self):
        s = self.scene
        self.ctx.upload_device(n_faces=s.n_faces, n_vertices=len(s.vertices), n_spheres=s.n_spheres, n_materials=len(s.materials), **self.ptr)

    def free(self):
        for p in self.ptr.values():
            if p:
                self.ctx.buffer_free(p)


def mixed_scene(seed=3, n_faces=700, n_spheres=300):
    from test_gpu_reference_mode import random_soup
    rng = np.random.default_rng(seed)
    scene = random_soup(rng, n_faces, n_spheres)
    mats = np.zeros(4, abi.MATERIAL_DTYPE)
    mats["kind"], mats["albedo"], mats["fuzz"], mats["ior"] = [0, 1, 2, 0], rng.uniform(0.2, 0.9, (4, 3)), [0, 0.2, 0, 0], [1, 1, 1.5, 1]
    return abi.SceneArrays(faces=scene.faces, vertices=scene.vertices, face_entity=scene.face_entity,
                           face_material=rng.integers(0, 4, scene.n_faces).astype(np.uint32), spheres=scene.spheres, sphere_color=scene.sphere_color,
                           sphere_entity=scene.sphere_entity, sphere_material=rng.integers(0, 4, scene.n_spheres).astype(np.uint32), materials=mats)


def test_upload_from_device_arrays_equals_upload_from_host_arrays(gpu_ctx):
    """Same kernels behind both entry points: frames, AOVs and path-traced frames (sweep and hierarchy) are identical."""
    import oraclelib as ol
    scene = mixed_scene()
    w, h = 120, 68
    cam = abi.reference_camera(w, h)
    pt = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=4, max_depth=8, seed=3)
    gpu_ctx.upload(scene)
    st = gpu_ctx.stats()
    assert st.upload_ms > 0 and st.upload_device_ms > 0 and st.h2d_ms > 0
    host = (gpu_ctx.render_aov(cam, abi.make_params(w, h)), gpu_ctx.render(cam, pt),
            gpu_ctx.render(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=4, max_depth=8, seed=3, flags=abi.FLAG_BVH)))
    d = DeviceScene(gpu_ctx, scene)
    try:
        d.upload()
        assert gpu_ctx.stats().h2d_ms == 0
        dev = (gpu_ctx.render_aov(cam, abi.make_params(w, h)), gpu_ctx.render(cam, pt),
               gpu_ctx.render(cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=4, max_depth=8, seed=3, flags=abi.FLAG_BVH)))
    finally:
        d.free()
    for a, b in zip(host[0], dev[0]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(host[1], dev[1]) and np.array_equal(host[2], dev[2]) and np.array_equal(host[1], host[2])
    oframe, oprim, _, ot = ol.oracle_reference(scene, cam, w, h)
    assert np.array_equal(dev[0][0], oframe) and np.array_equal(dev[0][1], oprim) and np.array_equal(dev[0][3].view(np.uint32), ot.view(np.uint32))
    cpu, _, _ = ol.oracle_pathtrace(scene, cam, pt)
    assert np.array_equal(dev[1], cpu)


def test_validation_happens_on_the_device_and_reports_the_first_error(gpu_ctx):
    scene = mixed_scene(seed=4, n_faces=50, n_spheres=20)
    bad = mixed_scene(seed=4, n_faces=50, n_spheres=20)
    bad.faces["v"][31, 1] = 10 ** 6
    bad.faces["v"][40, 0] = 10 ** 6
    with pytest.raises(abi.Rt3Error, match="face 31 references vertex out of range"):
        gpu_ctx.upload(bad)
    bad = mixed_scene(seed=4, n_faces=50, n_spheres=20)
    bad.sphere_material[7] = 9
    with pytest.raises(abi.Rt3Error, match="sphere 7: material 9 out of range"):
        gpu_ctx.upload(bad)
    bad = mixed_scene(seed=4, n_faces=50, n_spheres=20)
    bad.materials["kind"][2] = 5
    with pytest.raises(abi.Rt3Error, match="material 2: unknown kind 5"):
        gpu_ctx.upload(bad)
    with pytest.raises(abi.Rt3Error, match="rt3_scene_upload has not been called"):   # a failed upload leaves no scene behind
        gpu_ctx.render(abi.reference_camera(32, 18), abi.make_params(32, 18))
    gpu_ctx.upload(scene)
    gpu_ctx.render(abi.reference_camera(32, 18), abi.make_params(32, 18))


def test_tessellate_into_device_arrays_and_build_the_scene_there(gpu_ctx):
    """Spheres tessellated straight into the caller's device arrays (no trip to the host), next to host-made triangles;
    the result equals the host round trip, array for array and frame for frame."""
    via_host = gpu_ctx.tessellate_spheres(SPHERES, first_vertex=3)
    nf, nv = via_host.n_faces, len(via_host.vertices)
    tri_v = np.zeros(3, abi.VERTEX_DTYPE)
    tri_v["xyz"] = [(-1, -1, -2.5), (1, -1, -2.5), (0, 1, -2.5)]
    tri_f = np.zeros(1, abi.FACE_DTYPE)
    tri_f["v"], tri_f["normal"], tri_f["color"] = (0, 1, 2), (0, 0, 1), (0.3, 0.6, 0.9)
    faces_p, verts_p, ent_p = gpu_ctx.buffer_alloc((nf + 1) * 48), gpu_ctx.buffer_alloc((nv + 3) * 16), gpu_ctx.buffer_alloc((nf + 1) * 4)
    try:
        gpu_ctx.buffer_write(faces_p, tri_f)
        gpu_ctx.buffer_write(verts_p, tri_v)
        gpu_ctx.buffer_write(ent_p, np.array([99], np.uint32))
        gpu_ctx.tessellate_spheres_device(SPHERES, 3, 1, faces_p, verts_p, ent_p)
        faces = gpu_ctx.buffer_read(faces_p, np.zeros(nf + 1, abi.FACE_DTYPE))
        verts = gpu_ctx.buffer_read(verts_p, np.zeros(nv + 3, abi.VERTEX_DTYPE))
        ent = gpu_ctx.buffer_read(ent_p, np.zeros(nf + 1, np.uint32))
        assert faces[1:].tobytes() == via_host.faces.tobytes() and verts[3:].tobytes() == via_host.vertices.tobytes()
        assert np.array_equal(ent[1:], via_host.face_entity) and ent[0] == 99 and faces[:1].tobytes() == tri_f.tobytes()
        w, h = 160, 90
        cam = abi.reference_camera(w, h)
        gpu_ctx.upload_device(n_faces=nf + 1, n_vertices=nv + 3, faces=faces_p, vertices=verts_p, face_entity=ent_p)
        dev = gpu_ctx.render_aov(cam, abi.make_params(w, h))
        gpu_ctx.upload(abi.SceneArrays(faces=faces, vertices=verts, face_entity=ent))
        host = gpu_ctx.render_aov(cam, abi.make_params(w, h))
        for a, b in zip(dev, host):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert (dev[1] != abi.NO_HIT).sum() > 1000
    finally:
        for p in (faces_p, verts_p, ent_p):
            gpu_ctx.buffer_free(p)


def test_host_backend_builds_the_c3_mesh_on_the_device(built):
    """CudaRenderSettings::device_tessellation: the host backend assembles the scene in device memory (a triangle from the
    host next to a tessellated sphere) and renders the same frame as with the host flatten."""
    frames = []
    for device in (False, True):
        hs = hostlib.HostScene()
        hs.add_triangle((-3, -1, -4), (-1, -1, -4), (-2, 1, -4), (0, 1, 0))
        hs.add_sphere((0.5, 0, -3), 1.0, 40, 30, (1, 0, 0))
        hs.create_renderer(mode=abi.MODE_REFERENCE, device_tessellation=device)
        hs.prerender()
        flat = hs.renderer_flat()
        assert flat.n_faces == hs.flatten().n_faces
        frame, _, _ = hs.render(200, 113)
        frames.append((frame, flat.faces.tobytes(), flat.vertices.tobytes()))
    assert np.array_equal(frames[0][0], frames[1][0])
    assert frames[0][1] == frames[1][1] and frames[0][2] == frames[1][2]
