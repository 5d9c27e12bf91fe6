"""N > 1 host logic on CPU: world_size-2 `gloo` processes check that the (region, sample) sharding is a
partition in rank order and that the optional single-consumer gather reassembles the single-process batch."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from genvarloader_b200._dist import gather_rows, shard_bounds, shard_indices, sharded_batches  # noqa: E402


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_sharded_batches_cover_everything_once():
    world, bs, n = 4, 5, 93
    seen = []
    per_rank = [list(sharded_batches(n, bs, r, world, shuffle_seed=7)) for r in range(world)]
    assert len({len(x) for x in per_rank}) == 1
    for step in range(len(per_rank[0])):
        seen.append(np.concatenate([per_rank[r][step] for r in range(world)]))
    allidx = np.concatenate(seen)
    assert sorted(allidx.tolist()) == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a "batch" of ragged rows: row i has length 3 + i % 5 and holds the value i
        ds_idx = np.arange(11)
        mine = shard_indices(ds_idx, rank, world)
        lens = torch.tensor([3 + int(i) % 5 for i in mine], dtype=torch.int64)
        data = torch.cat([torch.full((int(l),), int(i), dtype=torch.uint8) for i, l in zip(mine, lens)]) if len(mine) else torch.empty(0, dtype=torch.uint8)
        off = torch.zeros(len(mine) + 1, dtype=torch.int64)
        torch.cumsum(lens, 0, out=off[1:])
        g_data, g_off = gather_rows(data, off, dst=0)
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks reduction
        if rank == 0:
            q.put((g_data.numpy().tolist(), g_off.numpy().tolist(), float(t.item())))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    data, off, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lens = [3 + i % 5 for i in range(11)]
    exp = [i for i, l in enumerate(lens) for _ in range(l)]
    assert data == exp
    assert off == np.concatenate([[0], np.cumsum(lens)]).tolist()
    assert tmax == 2.0
