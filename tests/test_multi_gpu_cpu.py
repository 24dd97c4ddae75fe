"""CPU, world_size 2, gloo: the multi-GPU host logic (frame sharding, interleaved tile-row bands + gather to rank 0).
Each rank 'renders' its share with the CPU oracle (test infrastructure) so that only the partition / exchange code of
heightmap-ray-marcher_b200/multi_gpu.py is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
import scenes as S


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, scene_name, height_override, result_path, channels=4):
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    for p in (str(root), str(root / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import hmrm_pkg
    import oracle_lib as O

    hmrm_pkg.load()
    from heightmap_ray_marcher_b200 import multi_gpu as MG

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = dict(S.SCENE_BY_NAME[scene_name])
    if height_override:
        scene["height"] = height_override
    hm, cm = S.build_maps(scene["maps"], O.synth_maps)
    heights = O.update_heightmap(hm, scene["lum"], scene["min_height"], scene["max_height"])
    fr = H.oracle_frame(O, scene)
    Hh, W = scene["height"], scene["width"]
    local = np.zeros((MG.padded_height(Hh), W, 4), dtype=np.uint8)
    for t in MG.owned_tile_rows(Hh, rank, world):
        O.render(fr, heights, cm, want_steps=False, rows=(t * 4, min(t * 4 + 4, Hh)), framebuf=local[:Hh])
    # RGB8 bands (the north star's exchange format): the RGBA8 rows without their constant alpha byte
    local = np.ascontiguousarray(local[..., :channels])
    full = MG.gather_interleaved_bands(torch.from_numpy(local), Hh, rank, world)
    if rank == 0:
        want, _, _ = O.render(fr, heights, cm, want_steps=False)
        np.save(result_path, np.array([int(np.array_equal(full.numpy(), want[..., :channels]))]))
    else:
        assert full is None
    # frame sharding: every frame has exactly one owner, owners rotate
    owners = torch.zeros(10, dtype=torch.int64)
    for n in MG.frames_of_rank(10, rank, world):
        owners[n] = rank + 1
    dist.all_reduce(owners)
    assert owners.tolist() == [(n % world) + 1 for n in range(10)]
    dist.destroy_process_group()


@pytest.mark.parametrize("scene_name,height,channels", [("persp_basic", None, 4), ("spher_basic", 178, 3), ("tiny_res", None, 3)])
def test_interleaved_band_gather_world2_gloo(scene_name, height, channels, tmp_path):
    port = _free_port()
    result = tmp_path / "ok.npy"
    mp.spawn(_worker, args=(2, port, scene_name, height, str(result), channels), nprocs=2, join=True)
    assert np.load(result)[0] == 1


def test_partition_helpers(hmrm):
    from heightmap_ray_marcher_b200 import multi_gpu as MG

    for height in (1, 4, 5, 178, 2160, 4320):
        for world in (1, 2, 3, 8):
            rows = sorted(r for k in range(world) for r in MG.owned_pixel_rows(height, k, world))
            assert rows == list(range(height))
    assert MG.padded_height(178) == 180 and MG.tile_rows(2160) == 540
    assert [MG.frame_owner(n, 4) for n in range(6)] == [0, 1, 2, 3, 0, 1]
