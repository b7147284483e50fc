"""Multi-GPU partitioning of the hot path: stereo pairs are independent, so N ranks (one process per GPU) each take a
contiguous block of the pair stream and run the whole path locally - weights replicated, NO data-path collective
(the reference does the same thing inside one process with nn.DataParallel's scatter/gather,
train.py:71 / predict.py:50).  The only communication is gathering the small disparity maps when a caller wants them on
one rank."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) block of `n_items` for `rank`; sizes differ by at most one, earlier ranks get the extra."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %r/%r" % (rank, world_size))
    base, extra = divmod(max(n_items, 0), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_pairs(left: torch.Tensor, right: torch.Tensor, rank: int, world_size: int):
    lo, hi = shard_range(left.shape[0], rank, world_size)
    return left[lo:hi], right[lo:hi]


def gather_disparities(local: torch.Tensor, n_total: int, rank: int, world_size: int, group=None,
                       dst: int = 0) -> Optional[torch.Tensor]:
    """Collects per-rank disparity blocks (possibly ragged / empty) on `dst` in pair order."""
    import torch.distributed as dist
    if world_size == 1:
        return local
    counts = [shard_range(n_total, r, world_size) for r in range(world_size)]
    max_n = max(hi - lo for lo, hi in counts)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world_size)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, counts)], dim=0)


def sharded_inference(forward_fn, left: torch.Tensor, right: torch.Tensor, rank: int, world_size: int, group=None):
    """Runs `forward_fn(left_block, right_block) -> disparity block` on this rank's pairs and gathers on rank 0."""
    l, r = shard_pairs(left, right, rank, world_size)
    if l.shape[0] > 0:
        out = forward_fn(l, r)
    else:
        h3, w3 = -(-left.shape[2] // 3), -(-left.shape[3] // 3)
        out = torch.zeros((0, 3 * h3, 3 * w3), dtype=torch.float32, device=left.device)
    return gather_disparities(out, left.shape[0], rank, world_size, group)
