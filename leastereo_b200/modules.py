"""Parameter containers with the reference's exact module tree, so ``state_dict`` keys, shapes, creation order
(and therefore seeded random init) are identical to ``retrain/LEAStereo.py`` and a checkpoint loads with
``strict=True``.

Only the 2D feature net (out of the hot path) computes with stock PyTorch ops here.  The 3D
containers (``ConvBR3d``, ``Cell3d``, ``newMatching``) hold parameters; their arithmetic is done by the CUDA
engine in ``engine.py`` - calling ``forward`` on them routes to that engine and fails loudly without it.

Reference citations: ``models/operations_3d.py:31-55`` (ConvBR), ``:84-90`` (Identity), ``:5-8`` (OPS);
``retrain/skip_model_3d.py:12-36`` (Cell ctor), ``:79-138`` (newMatching ctor), ``:176-185`` (freeze_layers);
``retrain/new_model_2d.py:78-165`` (newFeature); ``models/operations_2d.py:31-47`` (2D ConvBR).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .structure import build_cell_specs, scale_dimension


def _init_convbr(module: nn.Module) -> None:
    # kaiming-normal fan_out / relu on the conv, BN gamma=1 beta=0 (operations_3d.py:49-55).
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.Conv3d)):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class ConvBR3d(nn.Module):
    """Conv3d(bias=False) + BatchNorm3d + ReLU, flags ``bn``/``relu`` (operations_3d.py:31-47)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, bn=True, relu=True):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("the retrain path only instantiates stride-1 3D convs (SURVEY.md §0.3)")
        self.relu = relu
        self.use_bn = bn
        self.conv = nn.Conv3d(C_in, C_out, kernel_size, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm3d(C_out)
        _init_convbr(self)

    def forward(self, x):
        from .engine import conv_br_forward  # CUDA path; raises if the extension is missing
        return conv_br_forward(self, x)


class Identity3d(nn.Module):
    """``skip_connect`` at stride 1 (operations_3d.py:84-90)."""

    def forward(self, x):
        return x


def make_op_3d(is_conv: bool, C: int) -> nn.Module:
    return ConvBR3d(C, C, 3, 1, 1) if is_conv else Identity3d()


class Cell3d(nn.Module):
    """Parameter holder for one matching cell: ``pre_preprocess``, ``preprocess``, ``_ops`` (genotype row order)."""

    def __init__(self, spec):
        super().__init__()
        self.spec = spec
        self.C_out = spec.c_out
        self.C_prev = spec.c_prev
        self.C_prev_prev = spec.c_prev_prev
        self.downup_sample = spec.downup_sample
        if spec.downup_sample not in (-1, 0, 1):
            raise NotImplementedError("cell 0 on level >= 2 has no defined scale in the reference either")
        self.pre_preprocess = ConvBR3d(self.C_prev_prev, self.C_out, 1, 1, 0)
        self.preprocess = ConvBR3d(self.C_prev, self.C_out, 1, 1, 0)
        self._ops = nn.ModuleList(make_op_3d(c, self.C_out) for c in spec.op_is_conv)

    scale_dimension = staticmethod(scale_dimension)


class newMatching(nn.Module):
    """3D matching net, skip-connection variant (``retrain/skip_model_3d.py:78-174``)."""

    def __init__(self, network_path, cell_arch, args=None):
        super().__init__()
        self.args = args
        self._step = args.mat_step
        self._num_layers = args.mat_num_layers
        self._block_multiplier = args.mat_block_multiplier
        self._filter_multiplier = args.mat_filter_multiplier
        if self._num_layers != 12:
            # skip_model_3d.py:144-159 hard-codes cells[0..11] and the (1,4)/(4,8) skip pairs.
            raise NotImplementedError("the skip variant of the matching net is defined for exactly 12 cells")
        self.network_path = [int(v) for v in network_path]
        self.cell_arch_np = np.asarray(cell_arch)
        initial_fm = self._filter_multiplier * self._block_multiplier
        self.initial_fm = initial_fm

        self.stem0 = ConvBR3d(initial_fm * 2, initial_fm, 3, stride=1, padding=1)
        self.stem1 = ConvBR3d(initial_fm, initial_fm, 3, stride=1, padding=1)
        self.cell_specs = build_cell_specs(self.network_path, self.cell_arch_np, self._num_layers,
                                           self._filter_multiplier, self._block_multiplier, self._step)
        self.cells = nn.ModuleList(Cell3d(s) for s in self.cell_specs)
        self.last_3 = ConvBR3d(initial_fm, 1, 3, 1, 1, bn=False, relu=False)
        self.last_6 = ConvBR3d(initial_fm * 2, initial_fm, 1, 1, 0)
        self.last_12 = ConvBR3d(initial_fm * 4, initial_fm * 2, 1, 1, 0)
        self.last_24 = ConvBR3d(initial_fm * 8, initial_fm * 4, 1, 1, 0)
        self.conv1 = ConvBR3d(initial_fm * 4, initial_fm * 2, 3, 1, 1)
        self.conv2 = ConvBR3d(initial_fm * 4, initial_fm * 2, 3, 1, 1)

    def forward(self, x):
        """``x``: fp32 (B, 2C, D3, H3, W3) cost volume -> fp32 (B, 1, D3, H3, W3) matching cost."""
        from .engine import matching_forward
        return matching_forward(self, x)

    def freeze_layers(self, n_layers: int):
        print(f"Freezing first {n_layers} layers of matching net")
        frozen = [self.stem0, self.stem1]
        for i in range(min(n_layers, 12)):
            frozen.append(self.cells[i])
            if i == 4:
                frozen.append(self.conv1)
            elif i == 8:
                frozen.append(self.conv2)
        for mod in frozen:
            for p in mod.parameters():
                p.requires_grad = False


# ----------------------------------------------------------------------------------------------------------
# 2D feature net - the producer of the hot path's input.  Stock PyTorch ops (cuDNN on GPU); SURVEY.md §8(f).
# ----------------------------------------------------------------------------------------------------------

class ConvBR2d(nn.Module):
    def __init__(self, C_in, C_out, kernel_size, stride, padding, bn=True, relu=True):
        super().__init__()
        self.relu = relu
        self.use_bn = bn
        self.conv = nn.Conv2d(C_in, C_out, kernel_size, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(C_out)
        _init_convbr(self)

    def forward(self, x):
        x = self.conv(x)
        if self.use_bn:
            x = self.bn(x)
        if self.relu:
            x = F.relu(x, inplace=True)
        return x


class Identity2d(nn.Module):
    def forward(self, x):
        return x


class Cell2d(nn.Module):
    def __init__(self, spec, block_multiplier):
        super().__init__()
        self.spec = spec
        self.block_multiplier = block_multiplier
        self.C_out = spec.c_out
        self.pre_preprocess = ConvBR2d(spec.c_prev_prev, spec.c_out, 1, 1, 0)
        self.preprocess = ConvBR2d(spec.c_prev, spec.c_out, 1, 1, 0)
        self._ops = nn.ModuleList(ConvBR2d(spec.c_out, spec.c_out, 3, 1, 1) if c else Identity2d()
                                  for c in spec.op_is_conv)

    def forward(self, s0, s1):
        spec = self.spec
        prev_input = s1
        if spec.downup_sample != 0:
            size = [scale_dimension(s1.shape[2], spec.scale), scale_dimension(s1.shape[3], spec.scale)]
            s1 = F.interpolate(s1, size, mode="bilinear", align_corners=True)
        if s0.shape[2:] != s1.shape[2:]:
            s0 = F.interpolate(s0, tuple(s1.shape[2:]), mode="bilinear", align_corners=True)
        if s0.shape[1] != self.C_out:
            s0 = self.pre_preprocess(s0)
        s1 = self.preprocess(s1)
        states = [s0, s1]
        for step in spec.steps:
            states.append(sum(self._ops[k](states[j]) for j, k in step))
        return prev_input, torch.cat(states[-self.block_multiplier:], dim=1)


class newFeature(nn.Module):
    """2D feature net (``retrain/new_model_2d.py:78-165``): 3 stems (stride 3 at stem1), cells, 1/3-res head."""

    def __init__(self, network_path, cell_arch, args=None):
        super().__init__()
        self.args = args
        self._step = args.fea_step
        self._num_layers = args.fea_num_layers
        self._block_multiplier = args.fea_block_multiplier
        self._filter_multiplier = args.fea_filter_multiplier
        initial_fm = self._filter_multiplier * self._block_multiplier
        half_initial_fm = initial_fm // 2
        self.stem0 = ConvBR2d(3, half_initial_fm, 3, stride=1, padding=1)
        self.stem1 = ConvBR2d(half_initial_fm, initial_fm, 3, stride=3, padding=1)
        self.stem2 = ConvBR2d(initial_fm, initial_fm, 3, stride=1, padding=1)
        self.cell_specs = build_cell_specs([int(v) for v in network_path], np.asarray(cell_arch), self._num_layers,
                                           self._filter_multiplier, self._block_multiplier, self._step)
        self.cells = nn.ModuleList(Cell2d(s, self._block_multiplier) for s in self.cell_specs)
        self.last_3 = ConvBR2d(initial_fm, initial_fm, 1, 1, 0, bn=False, relu=False)
        self.last_6 = ConvBR2d(initial_fm * 2, initial_fm, 1, 1, 0)
        self.last_12 = ConvBR2d(initial_fm * 4, initial_fm * 2, 1, 1, 0)
        self.last_24 = ConvBR2d(initial_fm * 8, initial_fm * 4, 1, 1, 0)

    def forward(self, x):
        stem0 = self.stem0(x)
        stem1 = self.stem1(stem0)
        stem2 = self.stem2(stem1)
        out = (stem1, stem2)
        for cell in self.cells:
            out = cell(out[0], out[1])
        last = out[-1]
        h, w = stem2.shape[2], stem2.shape[3]

        def up(t, size):
            return F.interpolate(t, size=size, mode="bilinear", align_corners=True)

        if last.shape[2] == h:
            return self.last_3(last)
        if last.shape[2] == h // 2:
            return self.last_3(up(self.last_6(last), (h, w)))
        if last.shape[2] == h // 4:
            return self.last_3(up(self.last_6(up(self.last_12(last), (h // 2, w // 2))), (h, w)))
        if last.shape[2] == h // 8:
            t = up(self.last_24(last), (h // 4, w // 4))
            return self.last_3(up(self.last_6(up(self.last_12(t), (h // 2, w // 2))), (h, w)))
        raise RuntimeError("feature net ended on an unexpected level")
