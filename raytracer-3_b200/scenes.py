"""Synthetic scenes of the shapes BASELINE.json names.

Everything is derived from the reference's stateless hash RNG
(src/lib/shaders/random_v1.glsl:22-53) with fixed seeds, so the scenes are
identical on every box. Scenes are returned as ``abi.SceneArrays`` (the
flattened form ``rt3_scene_upload`` takes) plus the matching camera.
"""
import math

import numpy as np

from . import abi

_U32 = np.uint32


def hash1(x):
    """_random_hash, random_v1.glsl:22-29 (vectorised, wraps mod 2^32)."""
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x = (x + (x << 10)) & 0xFFFFFFFF
    x ^= x >> 6
    x = (x + (x << 3)) & 0xFFFFFFFF
    x ^= x >> 11
    x = (x + (x << 15)) & 0xFFFFFFFF
    return x.astype(np.uint32)


def uniform(counter, seed):
    """float in [0,1): _random_float_construct(hash(uvec2(counter, seed))), random_v1.glsl:31,38-53."""
    c = np.atleast_1d(np.asarray(counter, dtype=np.uint64))
    h = hash1(c ^ hash1(seed).astype(np.uint64))
    bits = (h & _U32(0x007FFFFF)) | _U32(0x3F800000)
    out = bits.astype(np.uint32).view(np.float32) - np.float32(1.0)
    return out if np.ndim(counter) else out[0]


def look_at_camera(look_from, look_at, vup, vfov_deg, aspect, aperture=0.0, focus_dist=1.0):
    """RTIOW look-at camera expressed as the reference Camera's four vectors."""
    f = np.float32
    theta = math.radians(vfov_deg)
    h = math.tan(theta / 2.0)
    vh = 2.0 * h
    vw = aspect * vh
    look_from, look_at, vup = (np.asarray(v, np.float64) for v in (look_from, look_at, vup))
    w = look_from - look_at
    w /= np.linalg.norm(w)
    u = np.cross(vup, w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    hor = focus_dist * vw * u
    ver = focus_dist * vh * v
    llc = look_from - hor / 2 - ver / 2 - focus_dist * w
    return abi.make_camera(look_from.astype(f), hor.astype(f), ver.astype(f), llc.astype(f),
                           lens_radius=aperture / 2.0, lens_u=u.astype(f), lens_v=v.astype(f))


def _materials(rows):
    m = np.zeros(len(rows), abi.MATERIAL_DTYPE)
    for i, (kind, albedo, fuzz, ior) in enumerate(rows):
        m[i]["kind"] = kind
        m[i]["albedo"] = albedo
        m[i]["fuzz"] = fuzz
        m[i]["ior"] = ior
    return m


def rtiow_four_spheres(width=400, height=225):
    """C1: the book-1 'default' scene (ground + three spheres), camera at the origin.

    Camera as the reference builds it: Camera::update(W, H, focal 1, viewport (W/H)*2 x 2).
    """
    spheres = np.array([[0, -100.5, -1, 100], [0, 0, -1, 0.5], [-1, 0, -1, 0.5], [1, 0, -1, 0.5]], np.float32)
    mats = _materials([
        (abi.MAT_LAMBERTIAN, (0.8, 0.8, 0.0), 0.0, 1.0),
        (abi.MAT_LAMBERTIAN, (0.1, 0.2, 0.5), 0.0, 1.0),
        (abi.MAT_DIELECTRIC, (1.0, 1.0, 1.0), 0.0, 1.5),
        (abi.MAT_METAL, (0.8, 0.6, 0.2), 0.0, 1.0),
    ])
    colors = np.array([m["albedo"] for m in mats], np.float32)
    scene = abi.SceneArrays(spheres=spheres, sphere_color=colors, sphere_material=np.arange(4, dtype=np.uint32),
                            sphere_entity=np.arange(4, dtype=np.uint32), materials=mats)
    cam = abi.reference_camera(width, height, focal_length=1.0, viewport_height=2.0)
    return scene, cam


def rtiow_cover(width=1200, height=800, seed=0x5EED):
    """C2 / C4: the book-1 cover scene (~485 spheres, mixed materials), look-from (13,2,3), vfov 20, aperture 0.1."""
    spheres = [[0, -1000, 0, 1000]]
    rows = [(abi.MAT_LAMBERTIAN, (0.5, 0.5, 0.5), 0.0, 1.0)]
    grid = 0
    for a in range(-11, 11):
        for b in range(-11, 11):
            xi = [float(uniform(grid * 8 + k, seed)) for k in range(8)]
            grid += 1
            cx, cz = a + 0.9 * xi[1], b + 0.9 * xi[2]
            if math.sqrt((cx - 4.0) ** 2 + (cz - 0.0) ** 2) <= 0.9:
                continue
            spheres.append([cx, 0.2, cz, 0.2])
            if xi[0] < 0.8:
                rows.append((abi.MAT_LAMBERTIAN, (xi[3] * xi[4], xi[5] * xi[6], xi[7] * xi[3]), 0.0, 1.0))
            elif xi[0] < 0.95:
                rows.append((abi.MAT_METAL, (0.5 + 0.5 * xi[3], 0.5 + 0.5 * xi[4], 0.5 + 0.5 * xi[5]), 0.5 * xi[6], 1.0))
            else:
                rows.append((abi.MAT_DIELECTRIC, (1.0, 1.0, 1.0), 0.0, 1.5))
    spheres += [[0, 1, 0, 1.0], [-4, 1, 0, 1.0], [4, 1, 0, 1.0]]
    rows += [(abi.MAT_DIELECTRIC, (1.0, 1.0, 1.0), 0.0, 1.5),
             (abi.MAT_LAMBERTIAN, (0.4, 0.2, 0.1), 0.0, 1.0),
             (abi.MAT_METAL, (0.7, 0.6, 0.5), 0.0, 1.0)]
    mats = _materials(rows)
    n = len(spheres)
    scene = abi.SceneArrays(spheres=np.array(spheres, np.float32), sphere_color=np.array([r[1] for r in rows], np.float32),
                            sphere_material=np.arange(n, dtype=np.uint32), sphere_entity=np.arange(n, dtype=np.uint32),
                            materials=mats)
    cam = look_at_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, width / height, aperture=0.1, focus_dist=10.0)
    return scene, cam


def random_spheres(n, seed=0xB200, width=1920, height=1080):
    """C5: n spheres, centres U[-100,100]^2 x U[-200,-5], radii U[0.05,0.5]; reference camera at the origin."""
    idx = np.arange(n, dtype=np.uint64)
    u = [uniform(idx * 8 + k, seed).astype(np.float64) for k in range(7)]
    spheres = np.stack([-100 + 200 * u[0], -100 + 200 * u[1], -200 + 195 * u[2], 0.05 + 0.45 * u[3]], axis=1).astype(np.float32)
    colors = np.stack([0.1 + 0.9 * u[4], 0.1 + 0.9 * u[5], 0.1 + 0.9 * u[6]], axis=1).astype(np.float32)
    scene = abi.SceneArrays(spheres=spheres, sphere_color=colors, sphere_entity=np.arange(n, dtype=np.uint32))
    cam = abi.reference_camera(width, height)
    return scene, cam
