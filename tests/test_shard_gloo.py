"""Multi-rank host logic on the CPU: world-size-2 gloo.  The state batch is partitioned
contiguously, every rank works on its slice alone, and the optional gather reassembles the
rows in order -- no other collective exists on the path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyjac_b200 import shard


def test_partition_covers_batch_contiguously():
    for n in (0, 1, 7, 8, 1020, 1 << 20):
        for world in (1, 2, 3, 8):
            parts = shard.partition(n, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
            assert all(a <= b for a, b in parts)
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) == (-(-n // world) if n else 0)


def _worker(rank, world, port, n, width, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        a, b = shard.my_slice(n, rank, world)
        # a rank's "evaluation" of its slice: a function of the global state index only
        idx = torch.arange(a, b, dtype=torch.float64)
        local = idx[:, None] * 10.0 + torch.arange(width, dtype=torch.float64)[None, :]
        full = shard.gather_rows(local, n, dst=0)
        if rank == 0:
            q.put(full.numpy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [9, 16, 1])
def test_two_rank_gather_reassembles_rows(n):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.arange(n)[:, None] * 10.0 + np.arange(3)[None, :]
    assert np.array_equal(full, want)
