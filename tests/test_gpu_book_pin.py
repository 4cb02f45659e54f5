"""Independent pin of the rows the reference does not implement, GPU half:
the CUDA path (through the C ABI) against the book's own program (oracle/rtiow_book.cpp: double precision, recursive
ray_color, rejection sampling, std::mt19937, abc-form sphere test with both roots) -- a checker that shares no code,
no random numbers and no formulation with the kernel. Tolerances: tests/bookpin.py (PSNR >= 40 dB per pixel at
equal spp, >= 46 dB on 4x4 blocks, mean radiance within 0.5 %, white furnace within 1 %)."""
import numpy as np
import pytest

import bookpin
from rt3_b200 import abi

pytestmark = pytest.mark.gpu


def gpu_render(ctx, flags=0):
    def render(scene, cam, params):
        params.flags |= flags
        ctx.upload(scene)
        frame = ctx.render(cam, params)
        return frame, ctx.read_radiance(params.width, params.height)
    return render


@pytest.mark.parametrize("name", list(bookpin.CASES))
def test_gpu_agrees_with_the_book(gpu_ctx, name):
    pix, blk, rel = bookpin.compare(name, gpu_render(gpu_ctx))
    print(f"{name}: {pix:.1f} dB / {blk:.1f} dB / {rel:.1e}")


@pytest.mark.parametrize("name", ["c1_default", "mesh_and_spheres"])
def test_gpu_hierarchy_agrees_with_the_book(gpu_ctx, name):
    bookpin.compare(name, gpu_render(gpu_ctx, abi.FLAG_BVH))


@pytest.mark.parametrize("name", list(bookpin.FURNACE))
def test_white_furnace(gpu_ctx, name):
    bookpin.furnace(name, gpu_render(gpu_ctx))
