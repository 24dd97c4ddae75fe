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
    """local_full: torch uint8 [padded_height, W, C] (C = 4 RGBA8 or 3 RGB8) with this rank's rows rendered.
    -> [Tmax, 4, W, C] packed."""
    import torch

    T = tile_rows(height)
    tmax = (T + world - 1) // world
    width, ch = local_full.shape[1], local_full.shape[2]
    view = local_full.view(T, TILE_ROWS, width, ch)
    own = view[rank::world]
    if own.shape[0] == tmax:
        return own.contiguous()
    out = torch.zeros((tmax, TILE_ROWS, width, ch), dtype=local_full.dtype, device=local_full.device)
    out[: own.shape[0]] = own
    return out


def gather_interleaved_bands(local_full, height: int, rank: int, world: int, dst: int = 0, group=None):
    """Gather every rank's interleaved tile rows (RGBA8 or RGB8 bands) to `dst`.  Returns the full [height, W, C] frame
    on dst, else None."""
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
    width, ch = local_full.shape[1], local_full.shape[2]
    full = torch.empty((T, TILE_ROWS, width, ch), dtype=local_full.dtype, device=local_full.device)
    for r in range(world):
        n = len(range(r, T, world))
        full[r::world] = bufs[r][:n]
    return full.view(T * TILE_ROWS, width, ch)[:height]


class _DeviceBytes:
    """CUDA array interface over a raw device pointer (so torch can view memory owned by libhmrm.so)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


PEER_CTRL_BYTES = 256        # HMRM_PEER_CTRL_BYTES


class PeerFrame:
    """One RGBA8 (channels = 4) or RGB8 (channels = 3) frame in the ROOT rank's device memory that every rank's kernel
    stores its tile rows into directly (CUDA IPC mapping + peer stores over NVLink / NVSwitch, include/hmrm.h "peer
    frames").  Replaces pack -> NCCL gather -> unpack of `gather_interleaved_bands`: the exchange is fused into the
    render kernel's own pixel stores.

    exchange = "stores" (default): the render kernel's own pixel stores go to the root's memory.  exchange = "copy": every
    non-root rank renders into a staging frame on its own device and pushes its tile rows with one strided
    device-to-device copy (large NVLink transfers; wins when many ranks write into one root).  The root always renders
    straight into its own frame.

    completion = "device" (default): the ranks tell the root "my bands are in" through words in the root's memory and
    the root tells them "buffer read" through another — stream memory operations executed by the GPU front ends (or
    one-thread kernels, HMRM_PEER_SYNC=kernels: csrc/peer_sync.cuh), stream-ordered, no collective, no host
    synchronisation.  completion = "allreduce": round 1's one-element all-reduce as the barrier
    (kept for the A/B and for backends without peer atomics).

    `buffers` frames rotate, so frame i + 1 may be rendered while the root still reads frame i:
        pf.render(frame, i, stream)      every rank (frame.band_count / band_index set by the caller)
        pf.complete(i, stream)           stream-ordered on the root: the frame is whole for what follows on `stream`
        pf.tensor(i)                     root: torch view of the frame
        pf.release(i, stream)            root: the buffer may be overwritten (use i + buffers)
    """

    def __init__(self, renderer, height: int, width: int, rank: int, world: int, device: int, root: int = 0,
                 buffers: int = 2, group=None, channels: int = 4, completion: str = "device", exchange: str = "stores"):
        import torch
        import torch.distributed as dist

        assert channels in (3, 4) and completion in ("device", "allreduce") and exchange in ("stores", "copy")
        assert exchange == "stores" or completion == "device"
        self.exchange = exchange
        self.r, self.height, self.width, self.channels = renderer, height, width, channels
        self.rank, self.world, self.root, self.group, self.device = rank, world, root, group, device
        self.completion = completion
        # NCCL: collectives are stream-ordered on the GPU.  gloo (tests: several ranks on ONE GPU, which NCCL does not
        # allow): host-side collectives after a stream synchronise.
        self.on_host = world > 1 and dist.get_backend(group) != "nccl"
        where = "cpu" if self.on_host else f"cuda:{device}"
        self.frame_bytes = (padded_height(height) * width * channels + 255) // 256 * 256
        nbytes = self.frame_bytes + PEER_CTRL_BYTES
        self.ptrs = []
        for _ in range(buffers):
            if rank == root:
                p = renderer.device_alloc(nbytes)          # zero-filled: arrived = released = error = 0
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
        self.stage = None
        if exchange == "copy" and rank != root:
            self.stage = [torch.zeros(self.frame_bytes, dtype=torch.uint8, device=f"cuda:{device}") for _ in range(buffers)]

    def pointer(self, i: int) -> int:
        """Device pointer (valid in THIS process) of buffer i mod buffers."""
        return self.ptrs[i % len(self.ptrs)]

    def ctrl(self, i: int) -> int:
        return self.pointer(i) + self.frame_bytes

    def use(self, i: int) -> int:
        return i // len(self.ptrs) + 1

    def render(self, frame, i: int, stream=None) -> None:
        """This rank's bands of frame i into the root's buffer (asynchronous, on `stream`)."""
        if self.stage is not None:
            self.r.render_peer_staged(frame, self.stage[i % len(self.stage)], self.pointer(i), self.ctrl(i), self.use(i), stream)
        elif self.completion == "device":
            self.r.render_peer(frame, self.pointer(i), self.ctrl(i), self.use(i), stream)
        else:
            self.r.render_device(frame, self.pointer(i), stream)

    def complete(self, i: int = 0, stream=None) -> None:
        """Stream-ordered: once this has run on the root, every rank's stores of frame i have landed in its buffer."""
        if self.world == 1 and self.completion != "device":
            return
        if self.completion == "device":
            if self.rank == self.root:
                self.r.peer_wait(self.ctrl(i), self.use(i), self.world, stream)
            return
        import torch
        import torch.distributed as dist

        if self.on_host:
            torch.cuda.current_stream().synchronize()
        dist.all_reduce(self.token, group=self.group)

    def release(self, i: int, stream=None) -> None:
        """Root: frame i has been read; its buffer may be overwritten by frame i + buffers."""
        if self.completion == "device" and self.rank == self.root:
            self.r.peer_release(self.ctrl(i), self.use(i), stream)

    def check(self) -> None:
        """Root: raise if a device-side wait timed out (a rank never arrived / the root never released)."""
        if self.completion == "device" and self.rank == self.root:
            for b in range(len(self.ptrs)):
                arrived, released, error = self.r.peer_status(self.ctrl(b))
                if error:
                    raise RuntimeError(f"peer frame buffer {b}: a device-side wait timed out "
                                       f"(arrived {arrived}, released {released})")

    def tensor(self, i: int):
        """Root only: torch uint8 view [height, W, channels] of buffer i mod buffers."""
        import torch

        if self.rank != self.root:
            return None
        full = torch.as_tensor(_DeviceBytes(self.pointer(i), (padded_height(self.height), self.width, self.channels)),
                               device=f"cuda:{self.device}")
        return full[: self.height]

    def close(self) -> None:
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)      # nobody unmaps / frees while another rank may still store or poll
        if self.rank != self.root:
            for p in self.ptrs:
                self.r.ipc_close(p)
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.rank == self.root:
            for p in self.ptrs:
                self.r.device_free(p)
        self.ptrs = []
