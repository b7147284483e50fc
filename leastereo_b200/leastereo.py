"""Drop-in ``LEAStereo`` module: same constructor, ``forward(left, right) -> disparity``, attributes and
``state_dict`` schema as the reference's ``retrain/LEAStereo.py:12-52``.

What differs is *how* the hot path runs: the cost volume (``retrain/LEAStereo.py:34-48``), the 3D matching net
(``retrain/skip_model_3d.py:140-174``) and the disparity head (``models/build_model_2d.py:45-57``) are executed
by hand-written sm_100a CUDA kernels through the C-ABI library (``include/leastereo_b200.h``); there is no
PyTorch/CPU fallback for them - without the compiled extension or off a CUDA device the call raises.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .modules import newFeature, newMatching


class DisparityRegression(nn.Module):
    """``models/build_model_2d.py:27-42``: sum_d p[b,d,h,w] * d.  Parameter-free."""

    def __init__(self, maxdisp, device=None):
        super().__init__()
        self.maxdisp = maxdisp
        self.device = device

    def forward(self, x):
        assert x.is_contiguous() is True
        from .engine import disparity_regression
        return disparity_regression(x, self.maxdisp)


class Disp(nn.Module):
    """``models/build_model_2d.py:45-57``: trilinear x3 upsample (align_corners=False) -> softmin over
    disparity -> regression, executed as ONE fused kernel that never writes the D x H x W volume."""

    def __init__(self, device=None, maxdisp=192):
        super().__init__()
        self.maxdisp = maxdisp
        self.softmax = nn.Softmin(dim=1)   # kept for attribute parity; parameter-free
        self.disparity = DisparityRegression(maxdisp=maxdisp, device=device)

    def forward(self, x):
        """``x``: fp32 (B, 1, D3, H3, W3) matching cost -> fp32 (B, 3*H3, 3*W3) disparity."""
        from .engine import disp_head_forward
        return disp_head_forward(x, self.maxdisp)


class LEAStereo(nn.Module):
    def __init__(self, args, device=None):
        super().__init__()
        network_path_fea, cell_arch_fea = np.load(args.net_arch_fea), np.load(args.cell_arch_fea)
        network_path_mat, cell_arch_mat = np.load(args.net_arch_mat), np.load(args.cell_arch_mat)
        print('Feature network path:{}\nMatching network path:{} \n'.format(network_path_fea, network_path_mat))

        self.maxdisp = args.maxdisp
        self.feature = newFeature(network_path_fea, cell_arch_fea, args=args)
        self.matching = newMatching(network_path_mat, cell_arch_mat, args=args)
        self.disp = Disp(device, self.maxdisp)
        self.use_cuda = getattr(args, "cuda", True)
        self.device = device
        # Engine knobs (not part of the reference API): operand planes of the split-precision layout
        # (2 = bf16x3, 3 = bf16x6 ~ fp32) and which conv kernel family runs ("tc" tcgen05, "simt" fp32 FMA).
        self.engine_options = {}

    def extract_features(self, x, y):
        """The two feature-net passes of ``retrain/LEAStereo.py:31-32`` (stock PyTorch, TF32 disabled so the
        producer stays fp32-exact like the CPU oracle - SURVEY.md §8(c) oracle hygiene)."""
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            if x.shape == y.shape and not self.training:
                # one batched call: BN is in eval mode so batching left+right is exact per sample
                f = self.feature(torch.cat([x, y], dim=0))
                return f[: x.shape[0]], f[x.shape[0]:]
            return self.feature(x), self.feature(y)

    def forward(self, x, y):
        from .engine import full_forward, hot_path_forward, _no_eval_autograd
        if not self.training:
            _no_eval_autograd(x, y)         # eval mode is inference-only: never hand back silently detached results
        if not self.training and x.is_cuda:
            out = full_forward(self, x, y)          # native feature net + hot path (engine option feature="native")
            if out is not None:
                return out
        fx, fy = self.extract_features(x, y)
        return hot_path_forward(self, fx, fy)

    def load_state_dict(self, state_dict, strict=True, **kw):
        # train.py saves the DataParallel wrapper's dict (keys carry "module."), predict.py strips it only on
        # CPU (predict.py:57-61); accept both.
        if state_dict and all(k.startswith("module.") for k in state_dict):
            state_dict = {k[len("module."):]: v for k, v in state_dict.items()}
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        from .engine import invalidate_cached_plans
        invalidate_cached_plans(self)
        return out
