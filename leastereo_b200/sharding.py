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


# ----------------------------------------------------------------------------------------------------------------
# training: batch-sharded data parallelism (train.py:71 uses nn.DataParallel; here one process per GPU)
# ----------------------------------------------------------------------------------------------------------------

def allreduce_gradients(parameters: Sequence[torch.nn.Parameter], world_size: int, group=None) -> int:
    """Averages the gradients of `parameters` across ranks with ONE all-reduce over a flat fp32 bucket
    (1.72 M elements = 6.9 MB for LEAStereo: latency-bound, so a single bucket beats per-tensor calls; over NVLink 5 /
    NVSwitch NCCL reduces it in a few tens of microseconds).  Parameters without a gradient on this rank (the
    reference's unused heads never get one) contribute zeros, so every rank reduces the same layout.
    BatchNorm statistics stay per replica, as under nn.DataParallel.  Returns the number of elements reduced."""
    import torch.distributed as dist
    params = [p for p in parameters if p.requires_grad]
    if world_size <= 1 or not params:
        return 0
    device = params[0].device
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=device)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is not None:
            flat[off: off + n].copy_(p.grad.reshape(-1))
        off += n
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world_size)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off: off + n].view_as(p.grad))
        off += n
    return int(flat.numel())
