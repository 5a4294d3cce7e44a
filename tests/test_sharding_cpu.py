"""CPU: clip sharding + output gather over a world_size-2 gloo group (the N>1 host logic of bench.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from axial_vs_b200 import sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = sharding.shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
            assert seen == list(range(n))
            sizes = sharding.shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_items):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.shard_range(n_items, rank, world)
    local = torch.arange(a, b, dtype=torch.float32)[:, None] * torch.ones(1, 3)
    full = sharding.gather_clip_outputs(local, n_items)
    assert full.shape == (n_items, 3)
    assert torch.equal(full[:, 0], torch.arange(n_items, dtype=torch.float32))
    dist.destroy_process_group()


def test_gather_world2_ragged():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 7), nprocs=2, join=True)
