"""The C-ABI library loads on a CPU-only box and exports every symbol include/rt3cuda.h declares.
No compute calls here: those are the -m gpu tests."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from rt3_b200 import abi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rt3cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt3_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(abi.CORE_LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in rt3cuda.h but not exported"
    assert sorted(abi.EXPORTED_SYMBOLS) == names


def test_headers_are_plain_c():
    """The boundary is a C ABI: both public headers must compile as C99 on their own (what a cgo / JNI / ctypes-generator would see)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    for name in ("rt3cuda.h", "rt3_rng.h"):
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", name)], check=True)


def test_struct_layouts_match_header():
    assert abi.FACE_DTYPE.itemsize == 48 and abi.VERTEX_DTYPE.itemsize == 16 and abi.MATERIAL_DTYPE.itemsize == 32
    assert C.sizeof(abi.Camera) == 4 * (12 + 1 + 6)
    assert C.sizeof(abi.Params) == 44
    assert C.sizeof(abi.Scene) == 8 + 4 * 8 + 8 + 5 * 8
    assert C.sizeof(abi.Stats) == 4 * 8 + 3 * 8 + 8 + 3 * 8 + 8 + 2 * 8 + 2 * 8   # ... + upload_ms, upload_device_ms + beam_rays, beam_tests


def py_partition_rows(h, tile, idx, cnt):
    if cnt <= 1:
        return h
    return sum(min(tile, h - t * tile) for t in range(idx, (h + tile - 1) // tile, cnt))


@pytest.mark.parametrize("h,tile,cnt", [(225, 8, 1), (225, 8, 2), (800, 8, 8), (2160, 16, 8), (7, 3, 4), (5, 8, 8), (36, 1, 3)])
def test_partition_rows_cover_the_image(built, h, tile, cnt):
    lib = abi.load_core()
    rows = [lib.rt3_partition_rows(h, tile, i, cnt) for i in range(cnt)]
    assert rows == [py_partition_rows(h, tile, i, cnt) for i in range(cnt)]
    assert sum(rows) == h


@pytest.mark.parametrize("h,tile,outer,inner", [(225, 8, 2, 3), (67, 4, 1, 3), (800, 2, 4, 2), (36, 1, 3, 5)])
def test_partitions_nest(built, h, tile, outer, inner):
    """A process that is part p of `outer` and splits its share over `inner` devices (CudaRenderer with several contexts) uses
    parts i * outer + p of outer * inner: together exactly the rows of part p."""
    from rt3_b200 import distributed
    lib = abi.load_core()
    for p in range(outer):
        mine = set(distributed.owned_rows(h, tile, p, outer).tolist())
        split = [distributed.owned_rows(h, tile, i * outer + p, outer * inner).tolist() for i in range(inner)]
        assert sum(len(x) for x in split) == len(mine) and set().union(*map(set, split)) == mine
        assert [len(x) for x in split] == [lib.rt3_partition_rows(h, tile, i * outer + p, outer * inner) for i in range(inner)]


def test_fails_loudly_without_a_device(built):
    """No CPU fallback: on a box without CUDA, creating a context is an error, not a slow path."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    lib = abi.load_core()
    handle = C.c_void_p()
    assert lib.rt3_create(C.byref(handle), 0) < 0
    assert b"no CUDA device" in lib.rt3_last_error() or b"CUDA" in lib.rt3_last_error()
    with pytest.raises(abi.Rt3Error):
        abi.Context(0)
