"""N>1 host-side logic on CPU: two gloo ranks render their row tiles (CPU oracle standing in for the
device) and the frame-end gather reassembles exactly the single-rank frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tile_rows, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import rt3_b200  # noqa: F401
    import oraclelib as ol
    from rt3_b200 import abi, distributed, scenes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 64, 37
    scene, cam = scenes.rtiow_four_spheres(w, h)
    params = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=4, max_depth=8, seed=11, tile_rows=tile_rows, part_index=rank, part_count=world)
    frame, _, rays = ol.oracle_pathtrace(scene, cam, params, n_threads=2)        # writes only the owned rows
    own = distributed.owned_rows(h, tile_rows, rank, world)
    assert len(own) == abi.load_core().rt3_partition_rows(h, tile_rows, rank, world)
    mask = np.zeros(h, bool)
    mask[own] = True
    assert (frame[~mask] == 0).all()
    t = torch.from_numpy(frame.view(np.int32).reshape(-1).copy())
    slab = distributed.pack_rows(t, w, h, tile_rows, rank, world)
    slabs = distributed.gather_slabs(dist, slab, rank, world)
    total = torch.tensor([rays], dtype=torch.int64)
    dist.all_reduce(total)
    if rank == 0:
        full = torch.zeros(h * w, dtype=torch.int32)
        for r in range(world):
            distributed.unpack_rows(slabs[r], full, w, h, tile_rows, r, world)
        single, _, single_rays = ol.oracle_pathtrace(scene, cam, abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=4, max_depth=8, seed=11), n_threads=2)
        ok = np.array_equal(full.numpy().view(np.uint32).reshape(h, w), single) and int(total.item()) == single_rays
        with open(out_path, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,tile_rows", [(2, 8), (2, 5), (3, 1)])
def test_gloo_ranks_reassemble_the_single_rank_frame(built, tmp_path, world, tile_rows):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(world, _free_port(), tile_rows, str(out)), nprocs=world, join=True)
    assert out.read_text() == "ok"


def test_owned_rows_partition_the_image():
    from rt3_b200 import distributed
    for h, tile, world in [(800, 8, 8), (225, 8, 2), (7, 3, 4), (2160, 16, 8)]:
        rows = np.concatenate([distributed.owned_rows(h, tile, r, world) for r in range(world)])
        assert sorted(rows.tolist()) == list(range(h))


class _NoSharing:
    """Stands in for a context whose GPU cannot share frames (no IPC / no peer access)."""

    def __init__(self, fail_on_owner):
        self.fail_on_owner = fail_on_owner

    def frame_alloc(self, n_pixels):
        if self.fail_on_owner:
            raise RuntimeError("cudaIpcGetMemHandle failed: operation not supported")
        raise AssertionError("this stand-in has no device memory")

    def frame_import(self, handle):
        raise RuntimeError("cudaIpcOpenMemHandle failed: peer access is not supported between these two devices")

    def frame_release(self, ptr):
        raise AssertionError("nothing was mapped")

    def frame_free(self, ptr):
        raise AssertionError("nothing was allocated")


def _shared_frame_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import rt3_b200  # noqa: F401
    from rt3_b200 import distributed
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shared = distributed.SharedFrame(_NoSharing(fail_on_owner=True), dist, 64 * 37, rank, world, torch.device("cpu"))
    verdict = f"ok={shared.ok} error={type(shared.error).__name__} ptr={shared.ptr}"
    shared.close()   # collective, must not touch the (absent) mapping
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(verdict)
    dist.barrier()
    dist.destroy_process_group()


def test_shared_frame_setup_fails_on_every_rank_together(built, tmp_path):
    """SharedFrame: when the owner cannot export its frame, every rank learns it in the same collective calls (no hang, no
    half-mapped state) and the caller can fall back to the NCCL gather."""
    mp.spawn(_shared_frame_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    for rank in range(3):
        assert (tmp_path / f"rank{rank}.txt").read_text() == "ok=False error=RuntimeError ptr=None"
