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


class _DeviceBytes:
    """CUDA array interface over a raw device pointer (so torch can view memory owned by libhmrm.so)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerFrame:
    """One RGBA8 frame in the ROOT rank's device memory that every rank's kernel stores its tile rows into directly
    (CUDA IPC mapping + peer stores over NVLink / NVSwitch, include/hmrm.h "peer frames").

    Replaces pack -> NCCL gather -> unpack of `gather_interleaved_bands`: the exchange is fused into the render
    kernel's own pixel stores, and the only collective left is a one-element all-reduce that tells the root the
    frame is complete.  `buffers` frames rotate so that the root may still be reading frame n while frame n+1 is
    being written (the root's read and its next barrier are ordered on one stream).
    """

    def __init__(self, renderer, height: int, width: int, rank: int, world: int, device: int, root: int = 0,
                 buffers: int = 2, group=None):
        import torch
        import torch.distributed as dist

        self.r, self.height, self.width = renderer, height, width
        self.rank, self.world, self.root, self.group, self.device = rank, world, root, group, device
        # NCCL: the completion barrier is stream-ordered on the GPU.  gloo (tests: several ranks on ONE GPU, which
        # NCCL does not allow): host-side barrier after a stream synchronise.
        self.on_host = world > 1 and dist.get_backend(group) != "nccl"
        where = "cpu" if self.on_host else f"cuda:{device}"
        nbytes = padded_height(height) * width * 4
        self.ptrs = []
        for _ in range(buffers):
            if rank == root:
                p = renderer.device_alloc(nbytes)
                handle = torch.tensor(list(renderer.ipc_export(p)), dtype=torch.uint8, device=where)
            else:
                p = 0
                handle = torch.empty(64, dtype=torch.uint8, device=where)
            if world > 1:
                dist.broadcast(handle, src=root, group=group)
            if rank != root:
                p = renderer.ipc_open(bytes(handle.cpu().tolist()))
            self.ptrs.append(p)
        self.token = torch.zeros(1, dtype=torch.float32, device=where)

    def pointer(self, i: int) -> int:
        """Device pointer (valid in THIS process) of buffer i mod buffers: pass it to Renderer.render_device."""
        return self.ptrs[i % len(self.ptrs)]

    def complete(self) -> None:
        """Stream-ordered: once this has run on the root, every rank's stores of the frame rendered before it on the
        current stream have landed in the root's buffer."""
        if self.world > 1:
            import torch
            import torch.distributed as dist

            if self.on_host:
                torch.cuda.current_stream().synchronize()
            dist.all_reduce(self.token, group=self.group)

    def tensor(self, i: int):
        """Root only: torch uint8 view [height, W, 4] of buffer i mod buffers."""
        import torch

        if self.rank != self.root:
            return None
        full = torch.as_tensor(_DeviceBytes(self.pointer(i), (padded_height(self.height), self.width, 4)),
                               device=f"cuda:{self.device}")
        return full[: self.height]

    def close(self) -> None:
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        if self.rank != self.root:
            for p in self.ptrs:
                self.r.ipc_close(p)
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.rank == self.root:
            for p in self.ptrs:
                self.r.device_free(p)
        self.ptrs = []
