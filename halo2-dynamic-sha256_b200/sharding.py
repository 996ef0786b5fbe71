"""Instance sharding across ranks: independent instances, contiguous ranges per rank, no data-path
collective.  After the hot path the ranks all_gather digests + checksums (64 B per instance)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_instances: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [first, last) of rank `rank` out of `world`: sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("bad rank")
    base, rem = divmod(n_instances, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def gather_results(digests, checksums, world: int):
    """all_gather of per-rank digests [n_r*D,32] u8 and checksums [n_r,4] i64 tensors (equal n_r on every rank, as in
    the weak-scaling bench) -> concatenated tensors in rank order + the job checksum (sum of all totals mod 2^64)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        allc, alld = [checksums], [digests]
    else:
        allc = [torch.empty_like(checksums) for _ in range(world)]
        alld = [torch.empty_like(digests) for _ in range(world)]
        dist.all_gather(allc, checksums)
        dist.all_gather(alld, digests)
    cks = torch.cat(allc, 0)
    job = 0
    for v in cks[:, 3].cpu().tolist():
        job = (job + (v & ((1 << 64) - 1))) & ((1 << 64) - 1)
    return torch.cat(alld, 0), cks, job
