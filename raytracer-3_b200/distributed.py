"""Frame-end gather for the row-tile partition (one process per GPU, torch.distributed plumbing).

Rendering needs no collective: tiles of ``tile_rows`` rows are dealt round-robin to the ranks
(the reference's own hook for this is the unused BlockInfo{x,y,w,h} uniform,
src/lib/shaders/raytracer/raytracer_v4.glsl:70-79) and every rank renders its tiles from a
replicated scene. Only the packed uint32 pixels travel, once per frame: each rank packs its
rows into a compact slab (padded to the largest slab so a plain gather works), rank 0 gathers
over NCCL (gloo in the CPU tests) and de-interleaves.
"""
import numpy as np
import torch


def owned_rows(height, tile_rows, part_index, part_count):
    """Global row indices rank ``part_index`` renders, top to bottom (== rt3_partition_rows ordering)."""
    rows = np.arange(height)
    return rows[(rows // max(tile_rows, 1)) % max(part_count, 1) == part_index]


def max_owned_rows(height, tile_rows, part_count):
    return max(len(owned_rows(height, tile_rows, r, part_count)) for r in range(part_count))


def pack_rows(frame, width, height, tile_rows, part_index, part_count, out=None):
    """Host/torch restatement of rt3_pack_partition: frame [H*W] -> slab [max_rows*W] (zero padded)."""
    rows = torch.as_tensor(owned_rows(height, tile_rows, part_index, part_count), device=frame.device)
    n = max_owned_rows(height, tile_rows, part_count) * width
    slab = out if out is not None else torch.zeros(n, dtype=frame.dtype, device=frame.device)
    slab[: len(rows) * width] = frame.view(height, width)[rows].reshape(-1)
    return slab


def unpack_rows(slab, frame, width, height, tile_rows, part_index, part_count):
    """Host/torch restatement of rt3_unpack_partition: slab -> the owner's rows of frame."""
    rows = torch.as_tensor(owned_rows(height, tile_rows, part_index, part_count), device=frame.device)
    frame.view(height, width)[rows] = slab[: len(rows) * width].view(len(rows), width)
    return frame


def gather_slabs(dist, slab, rank, world, dst=0):
    """Gathers every rank's (equal-sized) slab onto ``dst``; returns the list there, None elsewhere."""
    if world == 1:
        return [slab]
    bucket = [torch.empty_like(slab) for _ in range(world)] if rank == dst else None
    dist.gather(slab, bucket, dst=dst)
    return bucket
