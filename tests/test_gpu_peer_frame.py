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


@pytest.mark.parametrize("world,completion,channels", [(2, "device", 4), (3, "device", 3), (2, "allreduce", 4),
                                                       (3, "allreduce", 3)])
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
