"""Genotype / network-path decoding shared by the 2D feature net and the 3D matching net.

What the reference does (behaviour restated, not copied):
  * ``models/decoding_formulas.py:6-30`` ``network_layer_to_space``: level path -> one-hot (L,4,3) array
    ``[layer][level][down/same/up]``.
  * ``models/genotypes_3d.py:5-8`` / ``genotypes_2d.py``: two primitives, index 0 = skip, index 1 = 3x3(x3) conv.
  * ``retrain/skip_model_3d.py:93-124`` / ``retrain/new_model_2d.py:98-127``: per-cell
    ``(level, prev_level, prev_prev_level, downup_sample)`` and the channel multipliers.
  * ``retrain/skip_model_3d.py:54-72``: which state feeds which ``_ops[k]`` (ops are *typed* by genotype row
    order but *applied* in ascending branch order - SURVEY.md App. E quirk, preserved here).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

PRIMITIVES_3D = ["skip_connect", "3d_conv_3x3"]
PRIMITIVES_2D = ["skip_connect", "conv_3x3"]

_LEVEL_MULT = {0: 1, 1: 2, 2: 4, 3: 8}


def network_layer_to_space(net_arch: Sequence[int]) -> np.ndarray:
    """Level path ``[l0, l1, ...]`` -> float64 one-hot array of shape (L, 4, 3).

    Axis 2 is the sampling choice that reaches the layer: 0 = came from one level finer (down-sample),
    1 = same level, 2 = came from one level coarser (up-sample).  Layer 0 always uses slot 0.
    """
    path = [int(v) for v in net_arch]
    space = np.zeros((len(path), 4, 3))
    for i, level in enumerate(path):
        if i == 0:
            sample = 0
        else:
            delta = level - path[i - 1]
            if delta == 1:
                sample = 0
            elif delta == 0:
                sample = 1
            elif delta == -1:
                sample = 2
            else:
                raise ValueError("network path jumps more than one level at layer %d: %r" % (i, path))
        space[i, level, sample] = 1
    return space


@dataclass
class CellSpec:
    """Static description of one cell (2D or 3D)."""
    index: int
    level: int
    c_out: int               # filter_multiplier * {1,2,4,8}[level]
    c_prev_prev: int         # channels of s0 as it arrives
    c_prev: int              # channels of s1 as it arrives
    downup_sample: int       # -1 down, 0 none, +1 up (applied to s1)
    op_is_conv: List[bool]   # type of _ops[k], genotype row order
    # (state index fed, ops index used) per step, in application order
    steps: List[List[Tuple[int, int]]] = field(default_factory=list)

    @property
    def scale(self) -> float:
        return {-1: 0.5, 0: 1.0, 1: 2.0}[self.downup_sample]


def scale_dimension(dim: int, scale: float) -> int:
    """``Cell.scale_dimension`` (skip_model_3d.py:38-39): odd dims map (d-1)*s+1, even dims d*s."""
    dim = int(dim)
    if dim % 2 == 1:
        return int((float(dim) - 1.0) * scale + 1.0)
    return int(float(dim) * scale)


def cell_wiring(cell_arch: np.ndarray, steps: int) -> List[List[Tuple[int, int]]]:
    """For each step, the list of (state_index, ops_index) pairs that are summed into the new state."""
    branches = set(int(b) for b in np.asarray(cell_arch)[:, 0])
    n_states = 2
    offset = 0
    ops_index = 0
    wiring: List[List[Tuple[int, int]]] = []
    for _ in range(steps):
        this_step = []
        for j in range(n_states):
            if offset + j in branches:
                this_step.append((j, ops_index))
                ops_index += 1
        if not this_step:
            raise ValueError("cell step with no selected branch (sum of nothing)")
        wiring.append(this_step)
        offset += n_states
        n_states += 1
    return wiring


def build_cell_specs(network_path: Sequence[int], cell_arch: np.ndarray, num_layers: int,
                     filter_multiplier: int, block_multiplier: int, steps: int) -> List[CellSpec]:
    """Derive every cell's geometry the way the reference constructors do.

    Python's negative indexing in the reference (``network_arch[i-1]`` at i=0) only affects values that are
    then unused (cell 0 and 1 take their input widths from the stems), so it is not reproduced.
    """
    path = [int(v) for v in network_path]
    if len(path) < num_layers:
        raise ValueError("network path shorter than num_layers")
    space = network_layer_to_space(path)
    cell_arch = np.asarray(cell_arch)
    initial_fm = filter_multiplier * block_multiplier
    op_is_conv = [int(row[1]) == 1 for row in cell_arch]
    wiring = cell_wiring(cell_arch, steps)
    specs = []
    for i in range(num_layers):
        level = int(np.argmax(space[i].sum(axis=1)))
        c_out = filter_multiplier * _LEVEL_MULT[level]
        if i == 0:
            downup = -int(np.argmax(space[0].sum(axis=1)))
            c_pp, c_p = initial_fm, initial_fm
        else:
            downup = int(np.argmax(space[i].sum(axis=0))) - 1
            prev_level = path[i - 1]
            c_p = block_multiplier * filter_multiplier * _LEVEL_MULT[prev_level]
            if i == 1:
                c_pp = initial_fm
            else:
                c_pp = block_multiplier * filter_multiplier * _LEVEL_MULT[path[i - 2]]
        specs.append(CellSpec(index=i, level=level, c_out=c_out, c_prev_prev=c_pp, c_prev=c_p,
                              downup_sample=downup, op_is_conv=list(op_is_conv),
                              steps=[list(s) for s in wiring]))
    return specs
