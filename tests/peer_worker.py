"""Worker of tests/test_gpu_peer_frame.py: one rank of a 2- or 3-process job that shares ONE GPU.  Every rank renders
its interleaved tile rows of the same frame straight into rank 0's frame buffer (CUDA IPC); rank 0 compares the
result with the same frame rendered whole and writes the verdict to argv[1]."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import hmrm_pkg  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.cuda.set_device(0)
hmrm = hmrm_pkg.load()
from heightmap_ray_marcher_b200 import multi_gpu as MG  # noqa: E402

W, H = 322, 187            # not multiples of the 8x4 tile
r = hmrm.Renderer(0)
r.min_height, r.max_height = 0.0, 10.0
r.synth_maps(10, 1234)
stream = torch.cuda.Stream(device=0)
torch.cuda.set_stream(stream)
completion = sys.argv[2] if len(sys.argv) > 2 else "device"
channels = int(sys.argv[3]) if len(sys.argv) > 3 else 4
peer = MG.PeerFrame(r, H, W, rank, world, 0, channels=channels, completion=completion)
ok = []
for i, (proj, pos, vang) in enumerate([(1, (-3.0, 3.0, 14.0), 112.0), (2, (5.0, -5.0, 12.0), 120.0), (3, (-2.0, 2.0, 30.0), 125.0),
                                       (1, (12.0, 1.0, 11.0), 100.0)]):
    common = dict(projection=proj, screen_width=W, screen_height=H, cam_pos=pos, hang=hmrm.deg2rad(-45.0),
                  vang=hmrm.deg2rad(vang), hfov=hmrm.deg2rad(90.0), ortho_width=0.04, grid_width=0.01, step_dist=0.05)
    fmt = hmrm.PIXEL_RGB8 if channels == 3 else hmrm.PIXEL_RGBA8
    peer.render(r.frame(band_count=world, band_index=rank, pixel_format=fmt, **common), i, stream.cuda_stream)
    peer.complete(i, stream.cuda_stream)
    if rank == 0:
        got = peer.tensor(i).cpu().numpy()          # ordered after the completion on the current stream
        want = r.render(r.frame(**common))
        ok.append(bool(np.array_equal(got, want[..., :channels])) and (channels == 3 or bool((got[..., 3] == 255).all())))
    peer.release(i, stream.cuda_stream)
    if completion != "device":
        dist.barrier()      # nobody starts the next frame while rank 0 still reads this one (same buffer every 2 frames)
peer.check()
peer.close()
r.close()
if rank == 0:
    Path(sys.argv[1]).write_text(json.dumps(ok))
dist.destroy_process_group()
