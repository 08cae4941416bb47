"""The N > 1 host path on CPU: world_size-2 gloo processes shard frames round-robin and all-gather per-frame corner arrays
back into global frame order (rclc_b200/multi.py). No GPU compute here — the exchange is what is tested."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from rclc_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = multi.shard(n_frames, rank, world)
        # frame f's "corners": a recognisable 35 x 3 array
        local = np.stack([np.full((35, 3), float(f)) + np.arange(105).reshape(35, 3) * 1e-3 for f in mine]) if mine else np.zeros((0, 35, 3))
        out = multi.allgather_frames(local, n_frames, rank, world)
        q.put((rank, mine, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8, 1])
def test_shard_and_allgather_world2(n_frames):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    expect = np.stack([np.full((35, 3), float(f)) + np.arange(105).reshape(35, 3) * 1e-3 for f in range(n_frames)])
    owned = []
    for rank, mine, out in res:
        assert np.array_equal(out, expect), rank           # every rank holds every frame, in global order
        owned += mine
    assert sorted(owned) == list(range(n_frames))          # a partition of the frames


def test_shard_is_round_robin():
    from rclc_b200 import multi
    assert multi.shard(7, 0, 3) == [0, 3, 6] and multi.shard(7, 2, 3) == [2, 5] and multi.shard(2, 3, 4) == []
    assert np.array_equal(multi.allgather_frames(np.ones((4, 2)), 4, 0, 1), np.ones((4, 2)))
