"""GPU parity at the REAL shapes of all five BASELINE.json configurations, against the CPU oracle (not against the
other GPU path): full resolution, configured depth, the configured scene sizes (484 spheres, the 100 350-triangle mesh,
10^6 spheres), on row bands of the full-size frame -- and, where one row at the configured sample count is minutes of
brute force on the host, on windows of the configured samples (tests/fullsize.py explains the partition trick).
Frames and ray-segment counts must be equal bit for bit; the sweep and the hierarchy are both checked.
Reference loop: src/lib/renderer/SequentialRenderer.cpp:269-308 (per-pixel loop), :47-109 (closest hit)."""
import numpy as np
import pytest

import fullsize as fs
from rt3_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle_cache():
    return {}


def prepared(cache, name):
    """(scene, camera, oracle bands) of a configuration; the oracle runs once per configuration and module."""
    if name not in cache:
        c = fs.CONFIGS[name]
        scene, cam = c["scene"](c["w"], c["h"])
        cache[name] = (scene, cam, fs.oracle_bands(c, scene, cam))
    return cache[name]


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
@pytest.mark.parametrize("path", ["sweep", "bvh"])
def test_config_bands_equal_the_oracle(gpu_ctx, oracle_cache, name, path):
    c = fs.CONFIGS[name]
    scene, cam, cpu = prepared(oracle_cache, name)
    if name == "c5" and path == "sweep":
        # 10^6 spheres by brute force is what the hierarchy is for; the sweep still has to be right: one of the bands
        c = dict(c, rows=c["rows"][1:2])
    gpu_ctx.upload(scene)
    gpu = fs.gpu_bands(gpu_ctx, c, cam, abi.FLAG_BVH if path == "bvh" else 0)
    diff, rays_equal = fs.bands_match(gpu, {k: cpu[k] for k in gpu})
    assert rays_equal, f"{name}: ray-segment counts differ from the oracle on some band"
    assert diff == 0, f"{name}: {diff} of {len(gpu) * c['w']} band pixels differ from the oracle"


def test_full_frames_sweep_equals_hierarchy_and_bands_are_rows_of_them(gpu_ctx, oracle_cache):
    """C1 and C2 whole frames at the configured spp: hierarchy == sweep, and the band rows the oracle checked are rows of that frame."""
    for name in ("c1", "c2"):
        c = fs.CONFIGS[name]
        scene, cam, cpu = prepared(oracle_cache, name)
        gpu_ctx.upload(scene)
        sweep = gpu_ctx.render(cam, fs.params_for(c))
        tree = gpu_ctx.render(cam, fs.params_for(c, abi.FLAG_BVH))
        assert np.array_equal(sweep, tree)
        for (y, s, n), (pixels, _) in cpu.items():
            assert np.array_equal(sweep[y], pixels), f"{name}: row {y} of the full frame differs from the oracle"
