"""examples/c_abi_triangle.c: the boundary used from plain C (what a cgo / JNI / FFI binding would do).
CPU: it compiles as C99 against include/rt3cuda.h, links against librt3cuda.so and fails loudly without a device.
GPU: its frame hash is the compiled reference's (tests/golden/goldens.json)."""
import json
import os
import shutil
import subprocess

import pytest

from conftest import ROOT
from rt3_b200 import abi


def compile_example(tmp_path):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = str(tmp_path / "c_abi_triangle")
    libdir = os.path.dirname(abi.CORE_LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_triangle.c"), "-L", libdir, "-lrt3cuda", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_example_builds_and_has_no_cpu_path(built, tmp_path):
    exe = compile_example(tmp_path)
    if has_gpu():
        pytest.skip("a CUDA device is present: covered by the gpu test")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_example_renders_the_reference_frame(built, tmp_path):
    exe = compile_example(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "goldens.json")))["triangle_400x225"]
    fields = dict(zip(out.stdout.split()[0::2], out.stdout.split()[1::2]))
    assert fields["hash"] == meta["frame_fnv64_rows_0_to_Hm2"]
    assert fields["centre"] == meta["pixel_centre"] and int(fields["rays"]) == 400 * 225
