"""GPU: peer frames (include/hmrm.h, multi_gpu.PeerFrame).  Several processes — here all on cuda:0, with gloo for the
plumbing because NCCL refuses two ranks on one GPU — render their interleaved tile rows of one frame directly into
rank 0's frame buffer through a CUDA IPC mapping; the result must be byte-identical to the frame rendered whole.
On a multi-GPU box the same code path runs over NVLink with NCCL (bench.py --workload bands8k)."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

HERE = Path(__file__).resolve().parent


# completion = "allreduce" only: the device-side completion makes kernels of different ranks wait on one another, which
# is only safe when every rank has its own GPU (several processes spinning on ONE GPU can hit a context-switch
# timeout); its protocol is tested in one stream by test_device_side_completion_protocol_in_one_stream below and on
# real GPUs by bench.py --gpus N (which asserts the gathered frame).
@pytest.mark.parametrize("world,completion,channels", [(2, "allreduce", 4), (3, "allreduce", 3)])
def test_bands_stored_into_the_root_frame_equal_the_whole_frame(hmrm, world, completion, channels, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "verdict.json"
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(HERE / "peer_worker.py"), str(out), completion, str(channels)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=300)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), "\n".join(log[-1500:] for log in logs)
    assert json.loads(out.read_text()) == [True, True, True, True]


@pytest.mark.parametrize("sync", ["kernels", "memops"])
def test_device_side_completion_protocol_in_one_stream(hmrm, sync, monkeypatch):
    """hmrm_render_peer / hmrm_peer_wait / hmrm_peer_release (csrc/peer_sync.cuh) with three emulated ranks in ONE
    process and ONE stream: every wait is already satisfied by kernels earlier in the stream, so nothing spins on
    another launch.  Checks the counting (arrived = uses * ranks, released = uses), the frames (RGBA8 and RGB8, two
    rotating buffers, six frames) and that no wait timed out.  sync = memops: the same protocol through stream memory
    operations (HMRM_PEER_SYNC, read by hmrm_create): per-rank arrival words instead of the counter."""
    import numpy as np
    import torch

    from heightmap_ray_marcher_b200 import multi_gpu as MG

    W, H, ranks = 322, 187, 3
    monkeypatch.setenv("HMRM_PEER_SYNC", sync)
    r = hmrm.Renderer(0)
    try:
        r.min_height, r.max_height = 0.0, 10.0
        r.synth_maps(10, 1234)
        stream = torch.cuda.Stream(device=0)
        torch.cuda.set_stream(stream)
        for channels in (4, 3):
            fmt = hmrm.PIXEL_RGB8 if channels == 3 else hmrm.PIXEL_RGBA8
            pf = MG.PeerFrame(r, H, W, 0, 1, 0, channels=channels, completion="device")
            stage = torch.zeros(pf.frame_bytes, dtype=torch.uint8, device="cuda:0")
            for i in range(6):
                common = dict(projection=1 + i % 3, screen_width=W, screen_height=H, cam_pos=(-3.0 + i, 3.0, 14.0 + i),
                              hang=hmrm.deg2rad(-45.0), vang=hmrm.deg2rad(112.0), hfov=hmrm.deg2rad(90.0), ortho_width=0.04,
                              grid_width=0.01, step_dist=0.05)
                for rk in (2, 0, 1):         # any order: each call waits for the release of the previous use only
                    fr = r.frame(band_count=ranks, band_index=rk, pixel_format=fmt, **common)
                    if rk == 1:
                        # this "rank" exchanges through the copy engine: renders into its own staging frame, then pushes
                        # its tile rows with one strided device-to-device copy (hmrm_render_peer_staged)
                        r.render_peer_staged(fr, stage, pf.pointer(i), pf.ctrl(i), pf.use(i), stream.cuda_stream)
                    else:
                        r.render_peer(fr, pf.pointer(i), pf.ctrl(i), pf.use(i), stream.cuda_stream)
                r.peer_wait(pf.ctrl(i), pf.use(i), ranks, stream.cuda_stream)
                got = pf.tensor(i).cpu().numpy()
                r.peer_release(pf.ctrl(i), pf.use(i), stream.cuda_stream)
                want = r.render(r.frame(**common))
                assert np.array_equal(got, want[..., :channels]), (channels, i)
            stream.synchronize()
            for b in range(2):
                assert r.peer_status(pf.ctrl(b)) == (3 * ranks if sync == "kernels" else 0, 3, 0)
            pf.check()
            pf.close()
    finally:
        r.close()
