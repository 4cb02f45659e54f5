"""ctypes view of the C++ host backend (raytracer-3_b200/host/librt3host.so, host_capi.cpp)."""
import ctypes as C
import os

import numpy as np

from rt3_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "raytracer-3_b200", "host", "librt3host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(HOST_LIB)
        vp, fp, u32 = C.c_void_p, C.POINTER(C.c_float), C.c_uint32
        L.rt3host_last_error.restype = C.c_char_p
        L.rt3host_scene_create.restype = vp
        L.rt3host_scene_destroy.argtypes = [vp]
        L.rt3host_add_triangle.argtypes = [vp, fp, fp, fp, fp]
        L.rt3host_add_sphere.argtypes = [vp, fp, C.c_float, u32, u32, fp]
        L.rt3host_add_object.argtypes = [vp, C.c_char_p, fp, C.c_float, fp]
        L.rt3host_flatten.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), vp, vp, vp]
        L.rt3host_renderer_create.argtypes = [vp, C.c_int, u32, u32, u32, u32, u32, C.c_int]
        L.rt3host_renderer_create_multi.argtypes = [vp, C.POINTER(C.c_int), u32, u32, u32, u32, u32, u32, C.c_int, u32]
        L.rt3host_set_material.argtypes = [vp, u32, u32, fp, C.c_float, C.c_float]
        L.rt3host_prerender.argtypes = [vp]
        L.rt3host_render.argtypes = [vp, u32, u32, C.c_float, C.c_float, C.c_float, fp, vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.rt3host_camera_vectors.argtypes = [u32, u32, fp, fp]
        L.rt3host_render_progressive.argtypes = [vp, u32, u32, C.c_float, C.c_float, C.c_float, u32, vp]
        L.rt3host_add_scene_text.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(u32)]
        L.rt3host_radiance.argtypes = [vp, u32, u32, vp, C.c_char_p]
        L.rt3host_write_image.argtypes = [vp, u32, u32, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def _f(v):
    return (C.c_float * len(v))(*[float(x) for x in v])


class HostError(RuntimeError):
    pass


def write_image(frame, path, fmt):
    """Frame::to_ppm (fmt 'ppm') / Frame::to_png (fmt 'png') of the host backend on a [H, W] uint32 frame."""
    frame = np.ascontiguousarray(frame, np.uint32)
    if lib().rt3host_write_image(frame.ctypes.data, frame.shape[1], frame.shape[0], str(path).encode(), 1 if fmt == "png" else 0) != 0:
        raise HostError(lib().rt3host_last_error().decode())


class HostScene:
    """Entities created through ECS::create_*, rendered through RayTracer::CudaRenderer."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.rt3host_scene_create())

    def _ok(self, rc):
        if rc != 0:
            raise HostError(self.L.rt3host_last_error().decode())

    def add_triangle(self, p1, p2, p3, color):
        self._ok(self.L.rt3host_add_triangle(self.h, _f(p1), _f(p2), _f(p3), _f(color)))

    def add_sphere(self, center, radius, n_meridians, n_parallels, color):
        self._ok(self.L.rt3host_add_sphere(self.h, _f(center), radius, n_meridians, n_parallels, _f(color)))

    def add_object(self, path, center, scale, color):
        self._ok(self.L.rt3host_add_object(self.h, path.encode(), _f(center), scale, _f(color)))

    def add_scene_text(self, text, base_dir=""):
        """SceneLang (reference src/lib/sceneparser/SceneLang.md): appends the file's entities; returns (count, warnings)."""
        n = C.c_uint32()
        self._ok(self.L.rt3host_add_scene_text(self.h, text.encode(), str(base_dir).encode(), C.byref(n)))
        return n.value, [w for w in self.L.rt3host_last_error().decode().split("\n") if w]

    def flatten(self) -> abi.SceneArrays:
        nf, nv = C.c_uint32(), C.c_uint32()
        self._ok(self.L.rt3host_flatten(self.h, C.byref(nf), C.byref(nv), None, None, None))
        faces = np.zeros(nf.value, abi.FACE_DTYPE)
        verts = np.zeros(nv.value, abi.VERTEX_DTYPE)
        ent = np.zeros(nf.value, np.uint32)
        self._ok(self.L.rt3host_flatten(self.h, C.byref(nf), C.byref(nv), faces.ctypes.data, verts.ctypes.data, ent.ctypes.data))
        return abi.SceneArrays(faces=faces, vertices=verts, face_entity=ent)

    def renderer_flat(self) -> abi.SceneArrays:
        """The flattened scene the renderer holds after prerender() (read back from the device with device_tessellation)."""
        nf, nv = C.c_uint32(), C.c_uint32()
        self.L.rt3host_renderer_flat.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
        self._ok(self.L.rt3host_renderer_flat(self.h, C.byref(nf), C.byref(nv), None, None))
        faces, verts = np.zeros(nf.value, abi.FACE_DTYPE), np.zeros(nv.value, abi.VERTEX_DTYPE)
        self._ok(self.L.rt3host_renderer_flat(self.h, C.byref(nf), C.byref(nv), faces.ctypes.data, verts.ctypes.data))
        return abi.SceneArrays(faces=faces, vertices=verts)

    def create_renderer(self, device=0, mode=abi.MODE_REFERENCE, spp=1, max_depth=1, seed=1, flags=0, analytic_spheres=False,
                        device_tessellation=False):
        self._ok(self.L.rt3host_renderer_create(self.h, device, mode, spp, max_depth, seed, flags,
                                                int(analytic_spheres) | (2 if device_tessellation else 0)))

    def create_renderer_multi(self, devices, mode=abi.MODE_REFERENCE, spp=1, max_depth=1, seed=1, flags=0, analytic_spheres=False,
                              device_tessellation=False, tile_rows=0):
        arr = (C.c_int * len(devices))(*devices)
        self._ok(self.L.rt3host_renderer_create_multi(self.h, arr, len(devices), mode, spp, max_depth, seed, flags,
                                                      int(analytic_spheres) | (2 if device_tessellation else 0), tile_rows))

    def set_material(self, entity_index, kind, albedo=(1, 1, 1), fuzz=0.0, ior=1.5):
        self._ok(self.L.rt3host_set_material(self.h, entity_index, kind, _f(albedo), fuzz, ior))

    def prerender(self):
        self._ok(self.L.rt3host_prerender(self.h))

    def render(self, width, height, focal=2.0, vh=2.0, look=None):
        frame = np.zeros((height, width), np.uint32)
        ms, rays = C.c_double(), C.c_uint64()
        vw = float(np.float32(np.float32(width) / np.float32(height)) * np.float32(2.0))
        self._ok(self.L.rt3host_render(self.h, width, height, focal, vw, vh, _f(look) if look is not None else None, frame.ctypes.data,
                                       C.byref(ms), C.byref(rays)))
        return frame, ms.value, rays.value

    def render_progressive(self, width, height, passes, focal=2.0, vh=2.0):
        frames = np.zeros((passes, height, width), np.uint32)
        vw = float(np.float32(np.float32(width) / np.float32(height)) * np.float32(2.0))
        self._ok(self.L.rt3host_render_progressive(self.h, width, height, focal, vw, vh, passes, frames.ctypes.data))
        return frames

    def radiance(self, width, height, path=None):
        """CudaRenderer::read_radiance (and write_radiance_pfm when a path is given): (height, width, 3) float32."""
        rgb = np.zeros((height, width, 3), np.float32)
        self._ok(self.L.rt3host_radiance(self.h, width, height, rgb.ctypes.data, path.encode() if path else None))
        return rgb

    def close(self):
        if self.h:
            self.L.rt3host_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def camera_vectors(width, height, look):
    out = (C.c_float * 19)()
    assert lib().rt3host_camera_vectors(width, height, _f(look), out) == 0
    return np.array(list(out), np.float32)
