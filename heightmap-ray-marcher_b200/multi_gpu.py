"""Multi-GPU partitioning of the ray-march path (one process per GPU, torch.distributed for the plumbing).

Pixels are independent, so there is no collective on the data path except the final exchange:

* recording / flythrough (BASELINE configs[3]): frame n is rendered by rank n mod G — no collective at all;
* one big frame (BASELINE configs[4]): the 4-row tile rows of the frame are dealt round-robin to the ranks
  (`hmrm_frame.band_count / band_index`; interleaved rather than contiguous because sky rows cost nothing and
  terrain rows cost thousands of steps), and the RGBA8 rows are gathered to rank 0 (NCCL over NVLink on GPUs,
  gloo in the CPU tests).

The reference has no multi-GPU path (one OpenMP loop, main/hmap.cpp:978); this file is new surface.
"""
from __future__ import annotations

TILE_ROWS = 4      # screen tiles are 8x4 pixels


def frame_owner(frame_index: int, world: int) -> int:
    return frame_index % world


def frames_of_rank(n_frames: int, rank: int, world: int) -> range:
    return range(rank, n_frames, world)


def tile_rows(height: int) -> int:
    return (height + TILE_ROWS - 1) // TILE_ROWS


def owned_tile_rows(height: int, rank: int, world: int) -> range:
    return range(rank, tile_rows(height), world)


def owned_pixel_rows(height: int, rank: int, world: int) -> list[int]:
    rows = []
    for t in owned_tile_rows(height, rank, world):
        rows.extend(range(t * TILE_ROWS, min((t + 1) * TILE_ROWS, height)))
    return rows


def padded_height(height: int) -> int:
    return tile_rows(height) * TILE_ROWS


def pack_owned(local_full, height: int, rank: int, world: int):
    """local_full: torch uint8 [padded_height, W, 4] with this rank's rows rendered.  -> [Tmax, 4, W, 4] packed."""
    import torch

    T = tile_rows(height)
    tmax = (T + world - 1) // world
    view = local_full.view(T, TILE_ROWS, local_full.shape[1], 4)
    own = view[rank::world]
    if own.shape[0] == tmax:
        return own.contiguous()
    out = torch.zeros((tmax, TILE_ROWS, local_full.shape[1], 4), dtype=local_full.dtype, device=local_full.device)
    out[: own.shape[0]] = own
    return out


def gather_interleaved_bands(local_full, height: int, rank: int, world: int, dst: int = 0, group=None):
    """Gather every rank's interleaved tile rows to `dst`.  Returns the full [height, W, 4] frame on dst, else None."""
    import torch
    import torch.distributed as dist

    packed = pack_owned(local_full, height, rank, world)
    if world == 1:
        return local_full[:height]
    bufs = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    T = tile_rows(height)
    full = torch.empty((T, TILE_ROWS, local_full.shape[1], 4), dtype=local_full.dtype, device=local_full.device)
    for r in range(world):
        n = len(range(r, T, world))
        full[r::world] = bufs[r][:n]
    return full.view(T * TILE_ROWS, local_full.shape[1], 4)[:height]
