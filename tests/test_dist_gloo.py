"""CPU: world_size-2 gloo run of the sharding + pose-record gather (pix2pose_b200/dist.py)."""
import os
import socket
import subprocess
import sys

import numpy as np

from pix2pose_b200 import dist as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np
from pix2pose_b200 import dist as D
rank, local, world = D.init("gloo")
n = 11
idx = D.shard_indices(n, rank, world)
rec = np.zeros((len(idx), 16))
rec[:, 0] = idx * 10.0 + 1          # a per-detection payload derived from the global index
rec[:, 14] = 1
out = D.gather_records(rec, idx, n)
assert out.shape == (n, 16)
assert np.array_equal(out[:, 15], np.arange(n)), out[:, 15]
assert np.array_equal(out[:, 0], np.arange(n) * 10.0 + 1)
m = D.max_over_ranks(float(rank + 1))
assert m == world, m
D.barrier()
print("rank", rank, "ok")
"""


def test_shard_indices_partition():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 8):
            parts = [D.shard_indices(n, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if parts else np.array([])
            assert np.array_equal(allidx, np.arange(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_single_process_gather():
    rec = np.arange(3 * 16, dtype=np.float64).reshape(3, 16)
    out = D.gather_records(rec, [2, 0, 1], 3)
    assert np.array_equal(out[2, :15], rec[0, :15]) and np.array_equal(out[:, 15], [0, 1, 2])


def test_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()
