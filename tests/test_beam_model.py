"""CPU property test of the two beam inequalities behind the candidate lists of the primary rays (no GPU, no oracle):

  * sweep kernels, resident sphere scenes -- `beam_for_chunk`, csrc/rt3_kernels.cuh: a sphere is a candidate of a chunk of
    path items unless its centre is farther from the chunk's central ray than  Re + L' + s_hi k;
  * hierarchy kernels -- `beam_for_chunk_bvh`: a box is entered by the beam's walk unless the same holds for its bounding sphere
    (half diagonal + sqrt(3) ray_margin |o|max).

The formulas below restate the kernels' in numpy float32, in the same order. The property: for every primary ray the kernel can
generate for the chunk (`start_path`: jitter in [0, 1)^2, lens offset in the disc), the sphere its closest hit lies on -- computed
here by brute force in float64, with the margin the exact test's rounding can add -- is a candidate, and every box that contains
that sphere's (widened) box meets the beam. The GPU tests (`test_gpu_pathtrace.py::test_primary_candidate_lists_change_nothing`,
`test_gpu_bvh.py::test_primary_candidate_lists_through_the_hierarchy`) then require the frames to be the plain kernels' bit for bit.
"""
import math

import numpy as np
import pytest

from rt3_b200 import scenes

f32 = np.float32
SLACK = f32(3.814697265625e-06)      # RT3_FILTER_SLACK, csrc/rt3_device.cuh
SPHERE_MARGIN = f32(math.sqrt(3.814697265625e-06) * 1.0001)  # tree[1].ray_margin, csrc/rt3_core.cu build_bvh
TMIN = 0.001


def vec(c, name):
    return np.array(list(getattr(c, name))[:3], f32)


def beam_geometry(cam, W, H, xa, xb, y):
    """The set-up of beam_for_chunk / beam_for_chunk_bvh for the pixels [xa, xb] of row y (float32, the kernels' order)."""
    O, hor, ver, llc, lu, lv = (vec(cam, n) for n in ("origin", "horizontal", "vertical", "lower_left_corner", "lens_u", "lens_v"))
    iw, ih = f32(1) / f32(W - 1), f32(1) / f32(H - 1)
    u0, u1 = f32(xa) * iw, (f32(xb) + f32(1)) * iw
    v0, v1 = f32(H - 1 - y) * ih, (f32(H - 1 - y) + f32(1)) * ih
    um, vm = f32(0.5) * (u0 + u1), f32(0.5) * (v0 + v1)
    D = ((llc + um * hor) + vm * ver) - O
    n = lambda a: f32(np.sqrt(f32(a @ a)))  # noqa: E731
    len_hor, len_ver, len_o = n(hor), n(ver), n(O)
    tiny = f32(4.76837158203125e-07) * ((len_hor + len_ver) + (n(llc) + len_o))
    delta = (f32(0.5) * (u1 - u0) * len_hor + f32(0.5) * (v1 - v0) * len_ver) * f32(1.001) + tiny
    L = f32(cam.lens_radius)
    lens = L * (n(lu) + n(lv)) * f32(1.001) + tiny if L > 0 else tiny
    k = lens + delta
    dd = f32(D @ D)
    len_d = f32(np.sqrt(dd))
    ok = bool(len_d > f32(4) * k and len_d < 1e18 and k < 1e18)
    return dict(O=O, D=D, k=k, lens=lens, len_d=len_d, inv_dd=f32(1) / dd, inv_reach=f32(1) / (len_d - k), o_max=len_o + lens, tiny=tiny, ok=ok)


def misses(g, ctr, re):
    """True where the kernel drops a sphere / box with centre `ctr` [n, 3] and effective radius `re` [n] (float32)."""
    co = ctr - g["O"]
    proj = (co * g["D"]).sum(1, dtype=f32)
    co2 = (co * co).sum(1, dtype=f32)
    s_star = proj * g["inv_dd"]
    dist2 = co2 - s_star * proj
    s_hi = (np.maximum(s_star, f32(0)) * g["len_d"] + re + g["lens"]) * g["inv_reach"]
    reach = (re + g["lens"] + s_hi * g["k"]) * f32(1.0001)
    return dist2 > reach * reach + f32(1.9073486328125e-06) * co2


def sphere_candidates(g, sp):
    """beam_for_chunk: candidate mask over the spheres [n, 4]."""
    ctr, r2 = sp[:, :3], sp[:, 3] * sp[:, 3]
    o_slack = SLACK * (g["o_max"] * g["o_max"])
    re = np.sqrt((r2 + SLACK * ((ctr * ctr).sum(1, dtype=f32) + r2) + o_slack) * f32(1.0001), dtype=f32)
    return ~misses(g, ctr, re)


def boxes_met(g, lo, hi, ray_margin):
    """beam_for_chunk_bvh: which boxes [n, 3] x 2 the beam's walk enters."""
    grow = f32(1.7320508) * ray_margin * g["o_max"] * f32(1.0001) + g["tiny"]
    ctr = f32(0.5) * lo + f32(0.5) * hi
    half = f32(0.5) * hi - f32(0.5) * lo
    rho = np.sqrt((half * half).sum(1, dtype=f32), dtype=f32)
    re = (rho + grow) * f32(1.0001) + f32(4.76837158203125e-07) * (np.abs(ctr).sum(1, dtype=f32) + rho)
    return ~misses(g, ctr, re)


def primary_rays(cam, W, H, xa, xb, y, n, rng):
    """n rays of start_path for random pixels of [xa, xb] x {y}: (origins, unit directions), float64."""
    O, hor, ver, llc, lu, lv = (vec(cam, name).astype(np.float64) for name in ("origin", "horizontal", "vertical", "lower_left_corner", "lens_u", "lens_v"))
    x = rng.integers(xa, xb + 1, n)
    jx, jy = rng.random(n), rng.random(n)
    jx[: n // 4], jy[: n // 4] = rng.choice([0.0, 0.9999999], n // 4), rng.choice([0.0, 0.9999999], n // 4)  # the corners of the footprint
    u, v = (x + jx) / (W - 1), ((H - 1 - y) + jy) / (H - 1)
    org = np.repeat(O[None], n, 0)
    d = llc + u[:, None] * hor + v[:, None] * ver - O
    if cam.lens_radius > 0:
        rad, phi = np.sqrt(rng.random(n)), 2 * math.pi * rng.random(n)
        rad[: n // 4] = 1.0  # the rim of the lens
        off = cam.lens_radius * ((rad * np.cos(phi))[:, None] * lu + (rad * np.sin(phi))[:, None] * lv)
        org, d = org + off, d - off
    return org, d / np.linalg.norm(d, axis=1, keepdims=True)


def reportable(sp, o, d):
    """Spheres the exact test could report as the closest hit of ray (o, d): the closest geometric hit, plus every sphere whose
    hit is within the rounding of the float32 test of it (relative 1e-4 in t, discriminant slack as the filter's)."""
    c, r = sp[:, :3].astype(np.float64), sp[:, 3].astype(np.float64)
    oc = o - c
    h = oc @ d
    cc = (oc * oc).sum(1) - r * r
    disc = h * h - cc + float(SLACK) * ((c * c).sum(1) + r * r + o @ o)
    ok = disc >= 0
    sq = np.sqrt(np.where(ok, disc, 0.0))
    t1, t2 = -h - sq, -h + sq
    t = np.where(t1 >= TMIN * 0.999, t1, t2)
    ok &= t >= TMIN * 0.999
    if not ok.any():
        return np.zeros(0, int)
    t = np.where(ok, t, np.inf)
    best = t.min()
    return np.nonzero(t <= best * (1 + 1e-4) + 1e-6)[0]


CASES = [("cover, thin lens", lambda: scenes.rtiow_cover(192, 128), 192, 128),
         ("four spheres, pinhole", lambda: scenes.rtiow_four_spheres(160, 90), 160, 90),
         ("sphere cloud", lambda: scenes.random_spheres(40000, width=96, height=54), 96, 54)]


@pytest.mark.parametrize("name,make,W,H", CASES, ids=[c[0] for c in CASES])
def test_closest_hits_are_candidates_and_their_boxes_are_entered(name, make, W, H):
    scene, cam = make()
    sp = np.asarray(scene.spheres, f32)
    c64, r64 = sp[:, :3].astype(np.float64), sp[:, 3].astype(np.float64)
    # the boxes the hierarchy stores for spheres (csrc/rt3_upload.cuh build_spheres_kernel): centre +- sqrt(r^2 + slack (|c|^2 + r^2)), rounded outwards
    R = np.sqrt(r64 * r64 + float(SLACK) * ((c64 * c64).sum(1) + r64 * r64))
    lo = np.nextafter((c64 - R[:, None]).astype(f32), f32(-np.inf))
    hi = np.nextafter((c64 + R[:, None]).astype(f32), f32(np.inf))
    rng = np.random.default_rng(20261019)
    n_chunks = n_rays = n_hits = 0
    for _ in range(60):
        y = int(rng.integers(0, H))
        width = int(rng.choice([1, 1, 2, 4, 16]))
        xa = int(rng.integers(0, W - width + 1))
        xb = xa + width - 1
        g = beam_geometry(cam, W, H, xa, xb, y)
        assert g["ok"], "a regular camera must not be given up"
        cand = sphere_candidates(g, sp)
        met = boxes_met(g, lo, hi, SPHERE_MARGIN)
        n_chunks += 1
        org, dirs = primary_rays(cam, W, H, xa, xb, y, 48, rng)
        for o, d in zip(org, dirs):
            hits = reportable(sp, o, d)
            n_rays += 1
            n_hits += len(hits) > 0
            for j in hits:
                assert cand[j], f"{name}: sphere {j} can be hit by a primary ray of pixels [{xa}, {xb}] x {y} but is not a candidate"
                assert met[j], f"{name}: the beam's walk would not enter the box of sphere {j}"
                # any box that contains the sphere's box: unions with random other spheres (what the inner nodes are)
                others = rng.integers(0, len(sp), (6, 3))
                ulo = np.minimum(lo[j], lo[others].min(1))
                uhi = np.maximum(hi[j], hi[others].max(1))
                assert boxes_met(g, ulo, uhi, SPHERE_MARGIN).all(), f"{name}: the walk would skip an inner node above sphere {j}"
        # the lists stay short: that is what makes them worth having (the kernel gives up above 48 / 192)
        if width <= 2 and name != "sphere cloud":
            assert cand.sum() <= 48
    assert n_hits >= 100, "the test scene should be hit by a fair number of the rays"


def test_degenerate_beams_are_given_up_and_bad_data_is_kept():
    scene, cam = scenes.rtiow_cover(64, 40)
    sp = np.asarray(scene.spheres, f32).copy()
    g = beam_geometry(cam, 64, 40, 3, 3, 7)
    # NaN / infinite spheres and boxes stay in (the exact test decides)
    bad = np.array([[np.nan, 0, 0, 1], [0, 0, 0, np.nan], [np.inf, 0, 0, 1], [0, 0, 0, np.inf]], f32)
    with np.errstate(all="ignore"):
        assert sphere_candidates(g, bad).all()
        lo = np.array([[-np.inf] * 3, [np.nan] * 3, [0, 0, 0]], f32)
        hi = np.array([[np.inf] * 3, [np.nan] * 3, [np.inf, 1, 1]], f32)
        assert boxes_met(g, lo, hi, SPHERE_MARGIN).all()
    # a lens as large as the focus distance has no beam worth the name: ok = 0, the chunk's primary rays take the ordinary way
    wide = scenes.look_at_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.6, aperture=30.0, focus_dist=10.0)
    assert not beam_geometry(wide, 64, 40, 3, 3, 7)["ok"]
    # and a far-away sphere is dropped, a near one on the central ray kept
    D = g["D"] / np.linalg.norm(g["D"])
    on_axis = np.concatenate([g["O"] + f32(5) * D, [f32(0.1)]]).astype(f32)[None]
    side = np.cross(D, np.array([0, 1, 0], f32))
    off_axis = np.concatenate([g["O"] + f32(5) * D + f32(3) * side / np.linalg.norm(side), [f32(0.1)]]).astype(f32)[None]
    assert sphere_candidates(g, on_axis)[0] and not sphere_candidates(g, off_axis)[0]
