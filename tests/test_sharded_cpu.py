"""N > 1 host logic on CPU: world_size-2 `gloo` run of zk-apps_b200/sharded.py (point-range sharded MSM
with all_gather + local sum; round-robin proof batches with gather to rank 0).  The arithmetic backend
is the oracle here; tests/test_gpu_sharded.py runs the same classes with the GPU backend."""
import os
import socket

import zk_apps_b200  # noqa: F401
from zk_apps_b200 import sharded


def test_shard_range_and_round_robin():
    for n in (0, 1, 7, 8, 1 << 24, (1 << 24) + 5):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharded.round_robin(7, 1, 3) == [1, 4]
    assert sorted(i for r in range(8) for i in sharded.round_robin(1024, r, 8)) == list(range(1024))
    assert all(len(sharded.round_robin(1024, r, 8)) == 128 for r in range(8))


def test_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    from tests import _sharded_worker
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_sharded_worker.run, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
