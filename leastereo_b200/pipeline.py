"""Callers on either side of the hot path, on the device.

* input side  - ``predict.py:144-184`` / ``dataloaders/datasets/common.py:94-131``: per-channel z-normalisation and
  ``test_transform``'s pad / centre crop, fed by double-buffered pinned uploads of the 8-bit images;
* training    - the masked smooth-L1 loss of ``train.py:116-118,157`` (autograd function over two kernels), Adam over
  one flat parameter buffer (``train.py:76``) and the MultiStepLR schedule (``train.py:80``);
* evaluation  - EPE / 3-px error / bad-N of ``utils/metrics.py:6-46`` without a device->host copy of the disparity map.

Every arithmetic step runs in ``libleastereo_b200.so``; there is no PyTorch fallback (``kernels.get_ops`` raises when
the library is missing).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .kernels import Ops, get_ops


# ---------------------------------------------------------------------------------------------------------
# input pipeline
# ---------------------------------------------------------------------------------------------------------
def preprocess_pair(left_u8: torch.Tensor, right_u8: torch.Tensor, crop_h: int, crop_w: int,
                    ops: Optional[Ops] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Two uint8 (H, W, 3) device images -> the (1, 3, crop_h, crop_w) fp32 pair ``predict.py:test_transform`` builds."""
    ops = ops or get_ops()
    left = ops.normalize_pad_u8(left_u8, crop_h, crop_w)
    right = ops.normalize_pad_u8(right_u8, crop_h, crop_w)
    return left.unsqueeze(0), right.unsqueeze(0)


class InputPipeline:
    """Double-buffered host->device feed of 8-bit stereo pairs: while batch k is being normalised / run on the compute
    stream, batch k+1 is copied from pinned memory on a side stream.  ``submit`` takes (H, W, 3) uint8 CPU tensors (or
    numpy arrays) - or (batch, H, W, 3) stacks when ``batch > 1``; ``next`` returns the normalised
    (batch, 3, crop_h, crop_w) left / right tensors of the oldest submitted batch (``predict.py:144-184`` per pair).
    ``dst=(left, right)`` makes ``next`` write into caller-owned tensors (e.g. the static inputs of a CUDA graph).

    Only the 8-bit images cross PCIe: 2*H*W*3 bytes per pair, a quarter of the fp32 tensors the reference uploads."""

    def __init__(self, H: int, W: int, crop_h: int, crop_w: int, device, depth: int = 2, ops: Optional[Ops] = None,
                 batch: int = 1):
        self.ops = ops or get_ops()
        self.device = torch.device(device)
        self.crop = (int(crop_h), int(crop_w))
        self.depth, self.batch = depth, int(batch)
        B = self.batch
        self.host = [torch.empty((2, B, H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.dev = [torch.empty((2, B, H, W, 3), dtype=torch.uint8, device=self.device) for _ in range(depth)]
        self.out = [torch.empty((2, B, 3, crop_h, crop_w), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.sums = [torch.zeros((2, B, 6), dtype=torch.int64, device=self.device) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        self._head = self._tail = 0
        self.h2d_bytes_per_pair = 2 * H * W * 3
        self.h2d_bytes_per_batch = B * self.h2d_bytes_per_pair

    def submit(self, left_u8, right_u8):
        if self._head - self._tail >= self.depth:
            raise RuntimeError("InputPipeline: %d batches already in flight" % self.depth)
        s = self._head % self.depth
        if self._head >= self.depth:
            # the pinned slot still feeds the slot's previous (asynchronous) upload until ``ready[s]`` has completed:
            # wait for it on the HOST before overwriting the staging buffer (the caller need not synchronise)
            self.ready[s].synchronize()
        self.host[s][0].copy_(torch.as_tensor(left_u8).reshape(self.host[s][0].shape))
        self.host[s][1].copy_(torch.as_tensor(right_u8).reshape(self.host[s][1].shape))
        with torch.cuda.stream(self.copy_stream):
            if self._head >= self.depth:
                self.copy_stream.wait_event(self.consumed[s])        # the slot's previous batch has been normalised
            self.dev[s].copy_(self.host[s], non_blocking=True)
            self.ready[s].record(self.copy_stream)
        self._head += 1

    def next(self, dst: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._tail >= self._head:
            raise RuntimeError("InputPipeline: nothing submitted")
        s = self._tail % self.depth
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[s])
        outs = dst if dst is not None else (self.out[s][0], self.out[s][1])
        for k in range(2):
            for b in range(self.batch):
                self.ops.normalize_pad_u8(self.dev[s][k][b], self.crop[0], self.crop[1], out=outs[k][b],
                                          sums=self.sums[s][k][b])
        self.consumed[s].record(cur)
        self._tail += 1
        return outs[0], outs[1]


# ---------------------------------------------------------------------------------------------------------
# training step remainder
# ---------------------------------------------------------------------------------------------------------
class _MaskedSmoothL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, target, maxdisp):
        ops = get_ops()
        acc = ops.masked_smooth_l1(disp.detach(), target, maxdisp)
        ctx.save_for_backward(disp.detach(), target, acc)
        ctx.maxdisp = float(maxdisp)
        ctx.acc = acc
        return (acc[0] / acc[2].clamp_min(1.0)).to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        disp, target, acc = ctx.saved_tensors
        g = get_ops().masked_smooth_l1_bwd(disp, target, ctx.maxdisp, acc, 1.0)
        return g * gout, None, None


def masked_smooth_l1_loss(disp: torch.Tensor, target: torch.Tensor, maxdisp: float) -> torch.Tensor:
    """``F.smooth_l1_loss(disp[mask], target[mask], reduction='mean')`` with ``mask = (target < maxdisp) & (target >
    0.001)`` (``train.py:116-118,157``); 0 when no pixel is valid (the reference skips such batches, ``train.py:151``)."""
    return _MaskedSmoothL1.apply(disp, target, maxdisp)


def multistep_lr(base_lr: float, epoch: int, milestones: Sequence[int], gamma: float = 0.5) -> float:
    """``optim.lr_scheduler.MultiStepLR(optimizer, milestones, gamma=0.5)`` of ``train.py:80``."""
    return base_lr * gamma ** sum(1 for m in milestones if epoch >= m)


class FlatAdam:
    """``optim.Adam(model.parameters(), lr, betas=(0.9, 0.999))`` (``train.py:76``) as ONE kernel launch per step.

    The parameters are re-pointed into one flat fp32 buffer (their values, shapes, names and ``state_dict`` are
    unchanged) and so are their ``.grad`` tensors, which makes the flat gradient also the NCCL all-reduce bucket."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 ops: Optional[Ops] = None):
        self.ops = ops or get_ops()
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.grad[off:off + k].view_as(p)
                off += k
        self.lr, self.betas, self.eps, self.t = float(lr), (float(betas[0]), float(betas[1])), float(eps), 0

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p in self.params:                    # keep .grad pointing into the flat bucket (autograd accumulates in place)
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + k].view_as(p)
            off += k

    def _gather_grads(self):
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is not None and p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                self.grad[off:off + k].copy_(p.grad.reshape(-1))
            off += k

    def allreduce(self, world_size: int, group=None):
        """Average the gradients over the ranks: ONE all-reduce of the flat bucket, no packing copies (the parameters'
        ``.grad`` tensors are views of it).  Per-replica BatchNorm statistics, as under ``nn.DataParallel``."""
        if world_size <= 1:
            return
        import torch.distributed as dist
        self._gather_grads()
        dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
        self.grad.div_(world_size)

    def step(self):
        self._gather_grads()
        self.t += 1
        self.ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1],
                           self.eps, self.t)
        # the kernel wrote the parameters through a raw pointer (no autograd version bump): tell the inference plans
        from .engine import bump_param_generation
        bump_param_generation()


# ---------------------------------------------------------------------------------------------------------
# evaluation metrics
# ---------------------------------------------------------------------------------------------------------
def disparity_metrics(pred: torch.Tensor, target: torch.Tensor, maxdisp: float, thresholds=(1.0, 2.0, 3.0, 5.0),
                      ops: Optional[Ops] = None, float_diff: bool = False) -> Dict[str, float]:
    """The numbers ``evaluation.py:290-307`` prints for one frame, from one kernel and one 72-byte device->host read:
    ``epe`` (mean |d| over ``0.001 <= t <= maxdisp``, ``evaluation.py:290-292``), ``three_px_error`` and ``bad_N``
    (``utils/metrics.py:11-46`` - including its int64 truncation of |d|, so e.g. bad-1 counts ``|d| < 2`` as good) and
    ``train_epe`` (the strict-mask error ``train.py:203`` logs during validation).  ``float_diff=True`` gives the
    un-truncated variant, which is NOT what the reference reports."""
    ops = ops or get_ops()
    acc = ops.disparity_metrics(pred, target, maxdisp, thresholds, float_diff=float_diff).cpu().tolist()
    n = acc[0]
    nan = float("nan")
    out = {"valid": int(n), "epe": acc[1] / acc[7] if acc[7] > 0 else nan,
           "train_epe": acc[8] / n if n > 0 else nan,
           "three_px_error": 1.0 - acc[2] / n if n > 0 else nan}
    for k, t in enumerate(thresholds):
        out["bad_%g" % t] = 1.0 - acc[3 + k] / n if n > 0 else nan
    return out
