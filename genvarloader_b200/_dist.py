"""Multi-GPU plumbing: one process per GPU, every rank holds a full replica of the static tables and
works on a disjoint block of the flat (region, sample) index list.  There is NO
collective on the reconstruction path; `gather_rows` is the optional "single consumer" mode, an NCCL
gather over NVLink that is timed and reported separately from the HBM roofline.

`torch.distributed` is plumbing here: `nccl` on GPUs, `gloo` in the CPU tests (tests/test_dist_gloo.py)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n` items owned by `rank`: sizes differ by at most one, blocks are
    ordered by rank and cover 0..n exactly."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_indices(ds_idx, rank: int, world: int) -> np.ndarray:
    """This rank's block of a flat dataset-index list (row-major (region, sample) order is preserved, so
    concatenating the ranks' outputs in rank order reproduces the single-GPU batch)."""
    ds_idx = np.asarray(ds_idx)
    lo, hi = shard_bounds(len(ds_idx), rank, world)
    return ds_idx[lo:hi]


def sharded_batches(n_items: int, batch_size: int, rank: int, world: int, shuffle_seed: int | None = None):
    """Yield this rank's index batches for one pass over `n_items` (region, sample) pairs.  Every global
    batch of `batch_size * world` pairs is cut into `world` contiguous blocks (weak scaling: the per-GPU
    batch is fixed); a shared seed keeps the permutation identical on every rank without communication."""
    order = np.arange(n_items)
    if shuffle_seed is not None:
        order = np.random.default_rng(shuffle_seed).permutation(n_items)
    step = batch_size * world
    for s in range(0, n_items, step):
        chunk = order[s:s + step]
        yield shard_indices(chunk, rank, world)


def gather_rows(data: torch.Tensor, offsets: torch.Tensor, dst: int = 0, group=None):
    """Optional single-consumer mode: gather every rank's flat output rows on `dst` (rank order), returning
    `(data, offsets)` there and `(None, None)` elsewhere.  Rows are ragged across ranks, so sizes travel first.
    Bounded by the consumer's NVLink ingest (~0.8-0.9 TB/s), far below HBM: keep it off the hot path."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lens = offsets[1:] - offsets[:-1]
    meta = torch.tensor([data.shape[0], lens.numel()], dtype=torch.int64, device=data.device)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    sizes = [int(m[0]) for m in metas]
    nrows = [int(m[1]) for m in metas]
    if rank == dst:
        bufs = [torch.empty((sizes[r], *data.shape[1:]), dtype=data.dtype, device=data.device) for r in range(world)]
        lbufs = [torch.empty(nrows[r], dtype=lens.dtype, device=lens.device) for r in range(world)]
    else:
        bufs = lbufs = None
    # gather with point-to-point sends so ragged sizes need no padding
    if rank == dst:
        reqs = []
        for r in range(world):
            if r == dst:
                bufs[r].copy_(data)
                lbufs[r].copy_(lens)
            else:
                if sizes[r]:
                    reqs.append(dist.irecv(bufs[r], src=r, group=group))
                if nrows[r]:
                    reqs.append(dist.irecv(lbufs[r], src=r, group=group))
        for q in reqs:
            q.wait()
        all_lens = torch.cat(lbufs)
        out_off = torch.zeros(all_lens.numel() + 1, dtype=torch.int64, device=data.device)
        torch.cumsum(all_lens, 0, out=out_off[1:])
        return torch.cat(bufs), out_off
    reqs = []
    if data.shape[0]:
        reqs.append(dist.isend(data.contiguous(), dst=dst, group=group))
    if lens.numel():
        reqs.append(dist.isend(lens.contiguous(), dst=dst, group=group))
    for q in reqs:
        q.wait()
    return None, None
