"""Shared frame (rt3_frame_alloc / export / import, include/rt3cuda.h): the multi-GPU split without a gather.

One process owns the frame; another maps it over CUDA IPC and renders its row tiles straight into it.
A second GPU is not needed to exercise the path: the importing process here runs on the same device
(the mapping, the full-frame indexing and the ownership rules are the same; across GPUs the stores
travel over NVLink). The assembled frame must equal a one-partition render bit for bit.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from rt3_b200 import abi, scenes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, TILE_ROWS = 160, 96, 2

CHILD = r"""
import sys
sys.path.insert(0, {root!r})
import rt3_b200
from rt3_b200 import abi, scenes
handle = bytes.fromhex(sys.argv[1])
mode, part, parts = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
scene, cam = scenes.rtiow_cover({w}, {h})
ctx = abi.Context(0)
ctx.upload(scene)
frame = ctx.frame_import(handle)
params = abi.make_params({w}, {h}, mode=mode, spp=4, max_depth=8, seed=11, tile_rows={tile}, part_index=part, part_count=parts)
ctx.render_device(cam, params, frame, None)
ctx.stats()  # waits for the render
ctx.frame_release(frame)
ctx.close()
print("rows", ctx.lib.rt3_partition_rows({h}, {tile}, part, parts))
"""


def read_frame(ptr, n):
    import torch
    from rt3_b200.distributed import _DevicePointer
    torch.cuda.synchronize()
    return torch.as_tensor(_DevicePointer(ptr, n), device="cuda:0").cpu().numpy().view(np.uint32).reshape(H, W)


@pytest.mark.parametrize("mode", [abi.MODE_REFERENCE, abi.MODE_PATHTRACE])
def test_two_processes_render_into_one_frame(gpu_ctx, mode):
    scene, cam = scenes.rtiow_cover(W, H)
    gpu_ctx.upload(scene)
    whole = gpu_ctx.render(cam, abi.make_params(W, H, mode=mode, spp=4, max_depth=8, seed=11))
    frame = gpu_ctx.frame_alloc(W * H)
    try:
        handle = gpu_ctx.frame_export(frame)
        assert len(handle) == abi.IPC_HANDLE_BYTES
        parts = 3
        # partition 0 here, partitions 1 and 2 in other processes
        gpu_ctx.render_device(cam, abi.make_params(W, H, mode=mode, spp=4, max_depth=8, seed=11, tile_rows=TILE_ROWS, part_index=0, part_count=parts),
                              frame, None)
        gpu_ctx.stats()
        partial = read_frame(frame, W * H)
        own = (np.arange(H) // TILE_ROWS) % parts == 0
        assert np.array_equal(partial[own], whole[own])
        assert not partial[~own].any(), "rows of other partitions must stay untouched"
        code = CHILD.format(root=ROOT, w=W, h=H, tile=TILE_ROWS)
        for part in (1, 2):
            out = subprocess.run([sys.executable, "-c", code, handle.hex(), str(mode), str(part), str(parts)], capture_output=True, text=True, timeout=300)
            assert out.returncode == 0, out.stderr[-2000:]
        assert np.array_equal(read_frame(frame, W * H), whole)
    finally:
        gpu_ctx.frame_free(frame)


def test_argument_errors(gpu_ctx):
    with pytest.raises(abi.Rt3Error):
        gpu_ctx.frame_alloc(0)
    with pytest.raises(ValueError):
        gpu_ctx.frame_import(b"short")
    with pytest.raises(abi.Rt3Error):
        gpu_ctx.frame_import(bytes(abi.IPC_HANDLE_BYTES))  # not a handle anybody exported


def test_shared_frame_single_rank(gpu_ctx):
    """distributed.SharedFrame with one rank: the owner's torch view is the memory the kernels wrote."""
    import torch
    from rt3_b200 import distributed
    scene, cam = scenes.rtiow_four_spheres(W, H)
    gpu_ctx.upload(scene)
    params = abi.make_params(W, H, mode=abi.MODE_PATHTRACE, spp=2, max_depth=5, seed=3)
    whole = gpu_ctx.render(cam, params)
    shared = distributed.SharedFrame(gpu_ctx, None, W * H, 0, 1, torch.device("cuda", 0))
    try:
        gpu_ctx.render_device(cam, params, shared.ptr, None)
        shared.finish()
        gpu_ctx.stats()
        assert np.array_equal(shared.tensor.cpu().numpy().view(np.uint32).reshape(H, W), whole)
    finally:
        shared.close()
    assert shared.ptr is None
