"""CPU model of the hierarchy on BASELINE C5 (10^6 random spheres, primary rays from the origin): how many node
records does a ray read, and how much of that is owed to the slack the sphere boxes carry for the ill-conditioned
discriminant?  Runs here, no GPU: Morton order and splits as csrc/rt3_bvh.cuh builds them
(63-bit codes of the box centres, split at the highest differing bit), nearer child first, subtrees culled against
the closest hit so far.  Prints one JSON line per box variant.

Result (400 random pixels): boxes with the slack 126.2 visits and 2.41 leaf tests per ray (the GPU counts 129.5 and
2.43 over the whole frame), tight boxes 122.5 and 1.81 -- the slack costs 3 %, the rest is the scene: small spheres
two radii apart in a 200 x 200 x 195 box, rays that travel ~27 units before they hit something.
Usage: python profiles/c5_visits_sim.py [n_rays]"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rt3_b200  # noqa: F401,E402
from rt3_b200 import scenes  # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 400
scene, cam = scenes.random_spheres(1000000)
sp = scene.spheres.astype(np.float64)
c, r = sp[:, :3], sp[:, 3]
N = len(sp)


def spread21(v):
    x = v & 0x1fffff
    x = (x | x << 32) & 0x1f00000000ffff
    x = (x | x << 16) & 0x1f0000ff0000ff
    x = (x | x << 8) & 0x100f00f00f00f00f
    x = (x | x << 4) & 0x10c30c30c30c30c3
    x = (x | x << 2) & 0x1249249249249249
    return x


mn = c.min(0)
q = np.minimum(((c - mn) / (c.max(0) - mn) * 2097152).astype(np.uint64), 2097151)
key = spread21(q[:, 0]) | (spread21(q[:, 1]) << 1) | (spread21(q[:, 2]) << 2)
order = np.argsort(key, kind="stable")
keys, cs, rs = key[order], c[order], r[order]
splits = {}


def split(a, b):
    """Karras: the range [a, b) of sorted keys splits at the highest bit in which its first and last key differ."""
    if (a, b) not in splits:
        f, l = int(keys[a]), int(keys[b - 1])
        if f == l:
            splits[(a, b)] = (a + b) // 2
        else:
            bit = (f ^ l).bit_length() - 1
            splits[(a, b)] = a + int(np.searchsorted(keys[a:b], np.uint64(((f >> bit) | 1) << bit), "left"))
    return splits[(a, b)]


class Boxes:
    def __init__(self, radius):
        self.lo, self.hi, self.cache = cs - radius[order][:, None], cs + radius[order][:, None], {}

    def of(self, a, b):
        if (a, b) not in self.cache:
            self.cache[(a, b)] = (self.lo[a:b].min(0), self.hi[a:b].max(0))
        return self.cache[(a, b)]


def traverse(boxes, o, d):
    inv, best, visits, tests = 1.0 / d, math.inf, 0, 0
    stack = [(0, N, 0.0)]

    def slab(box):
        t0, t1 = (box[0] - o) * inv, (box[1] - o) * inv
        tin, tout = np.minimum(t0, t1).max(), np.maximum(t0, t1).min()
        return tin <= tout and tout >= 0 and tin <= best, tin

    while stack:
        a, b, tin = stack.pop()
        if tin > best:
            continue
        if b - a == 1:
            tests += 1
            oc = o - cs[a]
            h, cc = oc @ d, oc @ oc - rs[a] ** 2
            disc = h * h - cc
            if disc >= 0:
                t = -h - math.sqrt(disc)
                if t < 0.001:
                    t = -h + math.sqrt(disc)
                if 0.001 <= t < best:
                    best = t
            continue
        visits += 1
        m = split(a, b)
        (h0, t0), (h1, t1) = slab(boxes.of(a, m)), slab(boxes.of(m, b))
        if h0 and h1:
            near, far = ((a, m, t0), (m, b, t1)) if t0 <= t1 else ((m, b, t1), (a, m, t0))
            stack.append(far)
            stack.append(near)
        elif h0:
            stack.append((a, m, t0))
        elif h1:
            stack.append((m, b, t1))
    return visits, tests


W, H = 1920, 1080
hor, ver, llc = (np.array(list(v)[:3], float) for v in (cam.horizontal, cam.vertical, cam.lower_left_corner))
rng = np.random.default_rng(0)
rays = []
for _ in range(n_rays):
    x, y = rng.integers(0, W), rng.integers(0, H)
    d = llc + x / (W - 1) * hor + (H - 1 - y) / (H - 1) * ver
    rays.append(d / np.linalg.norm(d))
u = 2.0 ** -24
variants = {"with the discriminant's slack (R^2 = r^2 + 64u(|c|^2 + r^2), rt3_upload.cuh)": np.sqrt(r * r + 64 * u * ((c * c).sum(1) + r * r)),
            "tight (R = r)": r}
for name, radius in variants.items():
    boxes = Boxes(radius)
    v = t = 0
    for d in rays:
        a, b = traverse(boxes, np.zeros(3), d)
        v += a
        t += b
    print(json.dumps({"boxes": name, "rays": n_rays, "visits_per_ray": round(v / n_rays, 2), "tests_per_ray": round(t / n_rays, 2),
                      "mean_box_radius": round(float(radius.mean()), 4)}), flush=True)
