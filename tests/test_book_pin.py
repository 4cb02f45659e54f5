"""Independent pin of the rows the reference does not implement, CPU half:
the CPU restatement the GPU is compared with bit for bit (oracle/rt3_oracle.c orc_render_pathtrace) against the
book's own program (oracle/rtiow_book.cpp), which shares nothing with it. Tolerances: tests/bookpin.py.
The GPU half is tests/test_gpu_book_pin.py (same scenes, the CUDA path against the same checker)."""
import numpy as np
import pytest

import bookpin
import oraclelib as ol


def port_render(scene, cam, params):
    frame, accum, _ = ol.oracle_pathtrace(scene, cam, params, want_accum=True)
    return frame, (accum.astype(np.float64) / (params.spp * 2.0 ** 24)).astype(np.float32)


@pytest.mark.parametrize("name", ["c1_default", "inside_a_sphere", "hollow_glass", "mesh_and_spheres"])
def test_port_agrees_with_the_book(name):
    pix, blk, rel = bookpin.compare(name, port_render)
    print(f"{name}: {pix:.1f} dB / {blk:.1f} dB / {rel:.1e}")


@pytest.mark.parametrize("name", list(bookpin.FURNACE))
def test_white_furnace_port(name):
    bookpin.furnace(name, port_render)


@pytest.mark.parametrize("name", list(bookpin.FURNACE))
def test_white_furnace_book(name):
    """The checker itself conserves energy (otherwise agreeing with it would prove little)."""
    spheres, rows = bookpin.FURNACE[name]
    scene, cam = bookpin.sphere_scene(spheres, rows), bookpin.abi.reference_camera(64, 36, focal_length=1.0)
    rgb = ol.book_render(scene, cam, 64, 36, 64, 50, flags=bookpin.abi.FLAG_UNIFORM_SKY)
    assert rgb.max() <= 1.0 + 1e-6 and rgb.astype(np.float64).mean() >= 0.99
    if name != "default_scene_all_white":
        assert rgb.min() >= 1.0 - 1e-6
