"""CPU restatement (numpy / PyTorch-CPU) of the reference code on either side of the hot path - TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

Each function cites the reference lines it follows (paths relative to /root/reference).  The reference holds no tests or
golden vectors for this code, so the restatements are pinned against RUNS OF THE REFERENCE ITSELF:
tests/golden/make_io_golden.py imports utils/metrics.py and executes the source of predict.py's test_transform /
load_data (the module itself needs skimage / matplotlib, which are absent, so those two functions are exec'd from their
own source lines with PIL's Image.open replaced by an in-memory array) and of evaluation.py:290-292's EPE expression, and
commits inputs + outputs as tests/golden/io_*.npz; tests/test_oracle.py checks this module against those fixtures and,
where /root/reference is mounted, against the live functions.
"""
import numpy as np
import torch
import torch.nn.functional as F


def normalize_pair(left_u8: np.ndarray, right_u8: np.ndarray) -> np.ndarray:
    """predict.py:158-184 (load_data) == dataloaders/datasets/common.py:119-131: (6, H, W) float32, per-channel
    (x - mean) / std with numpy's fp64 mean and population std."""
    h, w = left_u8.shape[:2]
    out = np.zeros([6, h, w], "float32")
    for base, img in ((0, left_u8), (3, right_u8)):
        for c in range(3):
            ch = img[:, :, c]
            out[base + c, :, :] = (ch - np.mean(ch[:])) / np.std(ch[:])
    return out


def test_transform(temp_data: np.ndarray, crop_height: int, crop_width: int):
    """predict.py:144-156: pad bottom-right-aligned on zeros when the image fits, else centre crop."""
    _, h, w = np.shape(temp_data)
    if h <= crop_height and w <= crop_width:
        temp = temp_data
        temp_data = np.zeros([6, crop_height, crop_width], "float32")
        temp_data[:, crop_height - h: crop_height, crop_width - w: crop_width] = temp
    else:
        start_x = int((w - crop_width) / 2)
        start_y = int((h - crop_height) / 2)
        temp_data = temp_data[:, start_y: start_y + crop_height, start_x: start_x + crop_width]
    left = np.ones([1, 3, crop_height, crop_width], "float32")
    left[0, :, :, :] = temp_data[0: 3, :, :]
    right = np.ones([1, 3, crop_height, crop_width], "float32")
    right[0, :, :, :] = temp_data[3: 6, :, :]
    return left, right


def validity_mask(target: np.ndarray, max_disp):
    """utils/metrics.py:6-8 == train.py:116-118."""
    return (target < max_disp) & (target > 0.001)


def _abs_diff(pred: np.ndarray, true: np.ndarray, max_disp):
    """utils/metrics.py:13-17 / :40-43: `np.full(shape, 10000)` is an INT64 array, so the assignment truncates |d| toward
    zero; invalid pixels stand at 10000."""
    mask = validity_mask(true, max_disp)
    abs_diff = np.full(true.shape, 10000)
    abs_diff[mask] = np.abs(true[mask] - pred[mask])
    return mask, abs_diff


def three_px_error(pred: np.ndarray, true: np.ndarray, max_disp) -> float:
    """utils/metrics.py:11-21 (and :24-34, which also returns the `correct` map)."""
    mask, abs_diff = _abs_diff(pred, true, max_disp)
    correct = (abs_diff < 3) | (abs_diff < true * 0.05)
    return 1 - (float(np.sum(correct)) / float(len(np.argwhere(mask))))


def bad_pixel_frac(pred: np.ndarray, true: np.ndarray, max_disp, threshold) -> float:
    """utils/metrics.py:37-46."""
    mask, abs_diff = _abs_diff(pred, true, max_disp)
    correct = abs_diff <= threshold
    return 1 - (float(np.sum(correct)) / float(len(np.argwhere(mask))))


def epe(pred: np.ndarray, true: np.ndarray, max_disp) -> float:
    """evaluation.py:290-292: mean |prediction - disp| over the INCLUSIVE mask (disp >= 0.001) & (disp <= maxdisp)."""
    mask = np.logical_and(true >= 0.001, true <= max_disp)
    return float(np.mean(np.abs(pred[mask] - true[mask])))


def train_epe(pred: np.ndarray, true: np.ndarray, max_disp) -> float:
    """train.py:162 / :203: torch.mean(torch.abs(disp[mask] - target[mask])) over the strict validity mask."""
    mask = validity_mask(true, max_disp)
    return float(np.mean(np.abs(pred[mask] - true[mask])))


def masked_smooth_l1(disp: torch.Tensor, target: torch.Tensor, maxdisp) -> torch.Tensor:
    """train.py:116-118, :157."""
    mask = (target < maxdisp) & (target > 0.001)
    return F.smooth_l1_loss(disp[mask], target[mask], reduction="mean")
