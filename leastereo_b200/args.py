"""Model argument schema for the drop-in LEAStereo module.

Mirrors the reference's ``config_utils/leastereo_args.py:5-21`` (``LEAStereoArgs`` dataclass: the
``{fea,mat}_{num_layers,filter_multiplier,block_multiplier,step}`` integers plus four ``.npy`` architecture
paths).  ``maxdisp`` and ``cuda`` are attached by the reference's callers after parsing
(``retrain/LEAStereo.py:23,27`` read them), so they are plain optional fields here.
"""
from __future__ import annotations

import os
import tempfile
from dataclasses import dataclass
from typing import Optional

import numpy as np

# The searched architecture the reference ships in run/sceneflow/best/architecture/*.npy.
# They are four tiny integer arrays; embedding them lets the module be built with no files on disk.
SHIPPED_ARCH = {
    "feature_network_path": np.array([1, 0, 1, 0, 0, 0], dtype=np.uint8),
    "feature_genotype": np.array([[0, 1], [1, 0], [3, 1], [4, 1], [8, 1], [5, 1]], dtype=np.int64),
    "matching_network_path": np.array([1, 1, 2, 2, 1, 2, 2, 2, 1, 1, 0, 1], dtype=np.uint8),
    "matching_genotype": np.array([[1, 1], [0, 1], [3, 1], [4, 1], [8, 1], [6, 1]], dtype=np.int64),
}


@dataclass
class LEAStereoArgs:
    fea_num_layers: int = 6
    mat_num_layers: int = 12
    fea_filter_multiplier: int = 8
    mat_filter_multiplier: int = 8
    fea_block_multiplier: int = 4
    mat_block_multiplier: int = 4
    fea_step: int = 3
    mat_step: int = 3
    net_arch_fea: Optional[str] = None
    cell_arch_fea: Optional[str] = None
    net_arch_mat: Optional[str] = None
    cell_arch_mat: Optional[str] = None
    maxdisp: int = 192
    cuda: bool = True


def write_shipped_arch(directory: Optional[str] = None) -> dict:
    """Write the four shipped architecture arrays as ``.npy`` files and return their paths."""
    directory = directory or os.path.join(tempfile.gettempdir(), "leastereo_b200_arch")
    os.makedirs(directory, exist_ok=True)
    out = {}
    for name, arr in SHIPPED_ARCH.items():
        path = os.path.join(directory, name + ".npy")
        if not os.path.exists(path):
            tmp = path + ".%d.tmp.npy" % os.getpid()
            np.save(tmp, arr)
            os.replace(tmp, path)
        out[name] = path
    return out


def default_args(maxdisp: int = 192, cuda: bool = True, directory: Optional[str] = None) -> LEAStereoArgs:
    """``LEAStereoArgs`` pointing at the shipped SceneFlow architecture (predict_kitti15.sh:7-9 values)."""
    p = write_shipped_arch(directory)
    return LEAStereoArgs(
        net_arch_fea=p["feature_network_path"], cell_arch_fea=p["feature_genotype"],
        net_arch_mat=p["matching_network_path"], cell_arch_mat=p["matching_genotype"],
        maxdisp=maxdisp, cuda=cuda)
