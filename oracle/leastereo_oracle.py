"""CPU ORACLE for the LEAStereo hot path - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this file; nothing under ``leastereo_b200/`` does.  It is a restatement, in plain fp32 PyTorch-CPU /
numpy, of the reference's algorithm for

    cost volume          retrain/LEAStereo.py:34-48
    3D matching net      retrain/skip_model_3d.py:41-75 (Cell.forward), :140-174 (newMatching.forward)
    ConvBR               models/operations_3d.py:31-47
    disparity head       models/build_model_2d.py:27-57
    2D feature net       retrain/new_model_2d.py:41-75, :129-165 (the producer of the path's input)
    path decoding        models/decoding_formulas.py:6-30

written functionally over a ``state_dict`` (no nn.Module), so it shares no code with the product's module tree.
The arithmetic itself lives in a third-party dependency of the reference - PyTorch (ATen/oneDNN); the reference
pins torch==1.13.0 (freezed_cluster_requirements.txt:37), this image has torch 2.11.0 - so the oracle calls the
same ATen CPU operators the reference calls (conv3d, batch_norm, interpolate, softmin).

PINNING: the reference has no tests or golden vectors of its own.  The oracle is pinned
instead against outputs of the reference itself, produced by importing it in the build container:
``tests/golden/make_golden.py`` generates ``tests/golden/*.npz`` and ``tests/test_oracle.py`` checks this file
against them (bit-exact cost volume; disparity to 1e-4) and, when ``/root/reference`` is present, against the
live reference.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5

SHIPPED_ARCH = {
    # run/sceneflow/best/architecture/*.npy
    "feature_network_path": np.array([1, 0, 1, 0, 0, 0], dtype=np.uint8),
    "feature_genotype": np.array([[0, 1], [1, 0], [3, 1], [4, 1], [8, 1], [5, 1]], dtype=np.int64),
    "matching_network_path": np.array([1, 1, 2, 2, 1, 2, 2, 2, 1, 1, 0, 1], dtype=np.uint8),
    "matching_genotype": np.array([[1, 1], [0, 1], [3, 1], [4, 1], [8, 1], [6, 1]], dtype=np.int64),
}


# --------------------------------------------------------------------------------------------------------
# cost volume (retrain/LEAStereo.py:34-48)
# --------------------------------------------------------------------------------------------------------

def cost_volume_numpy(x: np.ndarray, y: np.ndarray, maxdisp: int) -> np.ndarray:
    """cost[b, c, d, h, w] = x[b,c,h,w] for w >= d ; cost[b, C+c, d, h, w] = y[b,c,h,w-d] for w >= d; else 0."""
    B, C, H, W = x.shape
    D = int(maxdisp / 3)
    cost = np.zeros((B, 2 * C, D, H, W), dtype=x.dtype)
    for d in range(D):
        if d >= W:
            break
        cost[:, :C, d, :, d:] = x[:, :, :, d:]
        cost[:, C:, d, :, d:] = y[:, :, :, : W - d]
    return cost


def cost_volume(x: torch.Tensor, y: torch.Tensor, maxdisp: int) -> torch.Tensor:
    return torch.from_numpy(cost_volume_numpy(x.detach().cpu().numpy(), y.detach().cpu().numpy(), maxdisp))


# --------------------------------------------------------------------------------------------------------
# ConvBR (models/operations_3d.py:31-47, operations_2d.py:31-47)
# --------------------------------------------------------------------------------------------------------

def conv_br(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, *, dims: int, stride: int = 1,
            bn: bool = True, relu: bool = True, training: bool = False,
            batch_stats: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    w = sd[prefix + ".conv.weight"]
    pad = (w.shape[-1] - 1) // 2
    conv = F.conv3d if dims == 3 else F.conv2d
    x = conv(x, w, None, stride, pad)
    if bn:
        if training:
            # batch statistics, biased variance for normalisation
            red = [0] + list(range(2, x.dim()))
            mean = x.mean(dim=red)
            var = x.var(dim=red, unbiased=False)
            if batch_stats is not None:
                n = x.numel() / x.shape[1]
                batch_stats[prefix] = (mean.clone(), var * (n / max(n - 1.0, 1.0)))
            # written out so that autograd differentiates through the batch statistics (train-mode BN)
            shape = [1, -1] + [1] * (x.dim() - 2)
            x = (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + BN_EPS) * sd[prefix + ".bn.weight"].view(shape) \
                + sd[prefix + ".bn.bias"].view(shape)
        else:
            mean, var = sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"]
            x = F.batch_norm(x, mean.clone(), var.clone(), sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"],
                             False, 0.0, BN_EPS)
    if relu:
        x = F.relu(x)
    return x


# --------------------------------------------------------------------------------------------------------
# cell wiring (retrain/skip_model_3d.py:13-75; identical logic in retrain/new_model_2d.py:12-75)
# --------------------------------------------------------------------------------------------------------

def scale_dimension(dim: int, scale: float) -> int:
    return int((float(dim) - 1.0) * scale + 1.0) if dim % 2 == 1 else int(float(dim) * scale)


def cell_geometry(path: Sequence[int], num_layers: int) -> List[int]:
    """downup_sample per cell: -level for cell 0 (skip_model_3d.py:105), level delta (-1/0/+1 as
    down/same/up => -1/0/+1) afterwards (:113-114 via network_layer_to_space, decoding_formulas.py:13-18)."""
    out = []
    for i in range(num_layers):
        if i == 0:
            out.append(-int(path[0]))
        else:
            delta = int(path[i]) - int(path[i - 1])
            out.append({1: -1, 0: 0, -1: 1}[delta])
    return out


def cell_forward(sd, prefix: str, cell_arch: np.ndarray, steps: int, block_multiplier: int, downup: int,
                 s0: torch.Tensor, s1: torch.Tensor, *, dims: int, training=False, batch_stats=None):
    mode = "trilinear" if dims == 3 else "bilinear"
    prev_input = s1
    c_out = sd[prefix + ".preprocess.conv.weight"].shape[0]
    if downup != 0:
        scale = 0.5 if downup == -1 else 2
        size = [scale_dimension(int(n), scale) for n in s1.shape[2:]]
        s1 = F.interpolate(s1, size, mode=mode, align_corners=True)
    if tuple(s0.shape[2:]) != tuple(s1.shape[2:]):
        s0 = F.interpolate(s0, tuple(s1.shape[2:]), mode=mode, align_corners=True)
    kw = dict(dims=dims, training=training, batch_stats=batch_stats)
    if s0.shape[1] != c_out:
        s0 = conv_br(sd, prefix + ".pre_preprocess", s0, **kw)
    s1 = conv_br(sd, prefix + ".preprocess", s1, **kw)

    selected = set(int(b) for b in cell_arch[:, 0])
    states = [s0, s1]
    offset = 0
    ops_index = 0
    for _ in range(steps):
        new_states = []
        for j, h in enumerate(states):
            if offset + j in selected:
                # _ops[k] is TYPED by genotype row k (skip_model_3d.py:33-36) but applied in branch order
                if int(cell_arch[ops_index][1]) == 1:
                    new_states.append(conv_br(sd, "%s._ops.%d" % (prefix, ops_index), h, **kw))
                else:
                    new_states.append(h)
                ops_index += 1
        s = new_states[0]
        for t in new_states[1:]:
            s = s + t
        offset += len(states)
        states.append(s)
    return prev_input, torch.cat(states[-block_multiplier:], dim=1)


# --------------------------------------------------------------------------------------------------------
# 3D matching net (retrain/skip_model_3d.py:140-174)
# --------------------------------------------------------------------------------------------------------

def matching_forward(sd, x: torch.Tensor, arch=SHIPPED_ARCH, *, steps=3, block_multiplier=4, num_layers=12,
                     prefix="matching", training=False, batch_stats=None, taps: dict = None) -> torch.Tensor:
    path = [int(v) for v in arch["matching_network_path"]]
    cell_arch = np.asarray(arch["matching_genotype"])
    downups = cell_geometry(path, num_layers)
    kw = dict(dims=3, training=training, batch_stats=batch_stats)

    def cell(i, a, b):
        out = cell_forward(sd, "%s.cells.%d" % (prefix, i), cell_arch, steps, block_multiplier, downups[i], a, b, **kw)
        if taps is not None:
            taps["cell%d" % i] = out[1]
        return out

    stem0 = conv_br(sd, prefix + ".stem0", x, **kw)
    stem1 = conv_br(sd, prefix + ".stem1", stem0, **kw)
    if taps is not None:
        taps["stem0"], taps["stem1"] = stem0, stem1
    out0 = cell(0, stem0, stem1)
    out1 = cell(1, out0[0], out0[1])
    out2 = cell(2, out1[0], out1[1])
    out3 = cell(3, out2[0], out2[1])
    out4 = cell(4, out3[0], out3[1])
    out4_cat = conv_br(sd, prefix + ".conv1", torch.cat((out1[1], out4[1]), 1), **kw)
    out5 = cell(5, out4[0], out4_cat)
    out6 = cell(6, out5[0], out5[1])
    out7 = cell(7, out6[0], out6[1])
    out8 = cell(8, out7[0], out7[1])
    out8_cat = conv_br(sd, prefix + ".conv2", torch.cat((out4[1], out8[1]), 1), **kw)
    out9 = cell(9, out8[0], out8_cat)
    out10 = cell(10, out9[0], out9[1])
    out11 = cell(11, out10[0], out10[1])
    last = out11[1]
    if taps is not None:
        taps["conv1"], taps["conv2"] = out4_cat, out8_cat

    d, h, w = x.shape[2:]

    def up(t, size):
        return F.interpolate(t, size=size, mode="trilinear", align_corners=True)

    if last.shape[3] == h:
        feat = last
    elif last.shape[3] == h // 2:
        feat = up(conv_br(sd, prefix + ".last_6", last, **kw), (d, h, w))
    elif last.shape[3] == h // 4:
        t = up(conv_br(sd, prefix + ".last_12", last, **kw), (d // 2, h // 2, w // 2))
        feat = up(conv_br(sd, prefix + ".last_6", t, **kw), (d, h, w))
    elif last.shape[3] == h // 8:
        t = up(conv_br(sd, prefix + ".last_24", last, **kw), (d // 4, h // 4, w // 4))
        t = up(conv_br(sd, prefix + ".last_12", t, **kw), (d // 2, h // 2, w // 2))
        feat = up(conv_br(sd, prefix + ".last_6", t, **kw), (d, h, w))
    else:
        raise RuntimeError("matching net ended on an unexpected level")
    return conv_br(sd, prefix + ".last_3", feat, dims=3, bn=False, relu=False)


# --------------------------------------------------------------------------------------------------------
# disparity head (models/build_model_2d.py:27-57)
# --------------------------------------------------------------------------------------------------------

def disp_head(mat: torch.Tensor, maxdisp: int) -> torch.Tensor:
    """(B,1,D3,H3,W3) -> (B,3*H3,3*W3): trilinear (align_corners=False) -> softmin(dim=1) -> sum_d p*d."""
    x = F.interpolate(mat, [maxdisp, mat.shape[3] * 3, mat.shape[4] * 3], mode="trilinear", align_corners=False)
    x = torch.squeeze(x, 1)
    p = F.softmin(x, dim=1)
    d = torch.arange(0, maxdisp, dtype=torch.float32, device=mat.device).reshape(1, maxdisp, 1, 1)
    return torch.sum(p * d, 1)


def disp_head_numpy_small(mat: np.ndarray, maxdisp: int) -> np.ndarray:
    """Independent float64 pure-numpy statement of the head for SMALL inputs (used to cross-check the ATen one)."""
    B, _, D3, H3, W3 = mat.shape

    def axis(out_n, in_n):
        scale = np.float32(in_n) / np.float32(out_n)
        i = np.arange(out_n, dtype=np.float32)
        src = np.maximum((i + np.float32(0.5)) * scale - np.float32(0.5), np.float32(0.0)).astype(np.float32)
        i0 = np.floor(src).astype(np.int64)
        i1 = np.minimum(i0 + 1, in_n - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float64)
        return i0, i1, 1.0 - l1, l1

    d0, d1, ad0, ad1 = axis(maxdisp, D3)
    h0, h1, ah0, ah1 = axis(3 * H3, H3)
    w0, w1, aw0, aw1 = axis(3 * W3, W3)
    v = mat[:, 0].astype(np.float64)
    v = v[:, d0] * ad0[None, :, None, None] + v[:, d1] * ad1[None, :, None, None]
    v = v[:, :, h0] * ah0[None, None, :, None] + v[:, :, h1] * ah1[None, None, :, None]
    v = v[:, :, :, w0] * aw0 + v[:, :, :, w1] * aw1
    v = -v
    v = v - v.max(axis=1, keepdims=True)
    e = np.exp(v)
    p = e / e.sum(axis=1, keepdims=True)
    return (p * np.arange(maxdisp, dtype=np.float64)[None, :, None, None]).sum(axis=1)


# --------------------------------------------------------------------------------------------------------
# 2D feature net (retrain/new_model_2d.py:129-165) - producer of the path's input
# --------------------------------------------------------------------------------------------------------

def feature_forward(sd, x: torch.Tensor, arch=SHIPPED_ARCH, *, steps=3, block_multiplier=4, num_layers=6,
                    prefix="feature", training=False, batch_stats=None) -> torch.Tensor:
    path = [int(v) for v in arch["feature_network_path"]]
    cell_arch = np.asarray(arch["feature_genotype"])
    downups = cell_geometry(path, num_layers)
    kw = dict(dims=2, training=training, batch_stats=batch_stats)
    stem0 = conv_br(sd, prefix + ".stem0", x, **kw)
    stem1 = conv_br(sd, prefix + ".stem1", stem0, stride=3, **kw)
    stem2 = conv_br(sd, prefix + ".stem2", stem1, **kw)
    out = (stem1, stem2)
    for i in range(num_layers):
        out = cell_forward(sd, "%s.cells.%d" % (prefix, i), cell_arch, steps, block_multiplier, downups[i],
                           out[0], out[1], **kw)
    last = out[1]
    h, w = stem2.shape[2:]

    def up(t, size):
        return F.interpolate(t, size=size, mode="bilinear", align_corners=True)

    if last.shape[2] == h:
        feat = last
    elif last.shape[2] == h // 2:
        feat = up(conv_br(sd, prefix + ".last_6", last, **kw), (h, w))
    elif last.shape[2] == h // 4:
        t = up(conv_br(sd, prefix + ".last_12", last, **kw), (h // 2, w // 2))
        feat = up(conv_br(sd, prefix + ".last_6", t, **kw), (h, w))
    else:
        t = up(conv_br(sd, prefix + ".last_24", last, **kw), (h // 4, w // 4))
        t = up(conv_br(sd, prefix + ".last_12", t, **kw), (h // 2, w // 2))
        feat = up(conv_br(sd, prefix + ".last_6", t, **kw), (h, w))
    return conv_br(sd, prefix + ".last_3", feat, dims=2, bn=False, relu=False)


# --------------------------------------------------------------------------------------------------------
# whole model (retrain/LEAStereo.py:30-52)
# --------------------------------------------------------------------------------------------------------

def leastereo_forward(sd, left: torch.Tensor, right: torch.Tensor, maxdisp: int, arch=SHIPPED_ARCH, *,
                      training=False, batch_stats=None, stages: dict = None) -> torch.Tensor:
    with torch.no_grad():
        fx = feature_forward(sd, left, arch, training=training, batch_stats=batch_stats)
        fy = feature_forward(sd, right, arch, training=training, batch_stats=batch_stats)
        # on a CUDA device (bench.py's gpu_baseline leg) the slice-assignment statement of LEAStereo.py:34-48 is used
        cost = cost_volume(fx, fy, maxdisp) if not fx.is_cuda else cost_volume_torch(fx, fy, maxdisp)
        mat = matching_forward(sd, cost, arch, training=training, batch_stats=batch_stats)
        disp = disp_head(mat, maxdisp)
    if stages is not None:
        stages.update(fx=fx, fy=fy, cost=cost, mat=mat, disp=disp)
    return disp


def calibrate_bn(sd, left, right, maxdisp, arch=SHIPPED_ARCH) -> Dict[str, torch.Tensor]:
    """"Calibrated" parity regime of SURVEY.md §8(d): one fp32 train-mode pass with BN momentum 1.0 sets the
    running statistics to the batch statistics; returns a new state_dict (weights untouched).

    The reference runs the feature net twice per forward (left then right, LEAStereo.py:31-32), so with momentum
    1.0 the RIGHT image's statistics are the ones that survive for ``feature.*`` - reproduced here.
    """
    stats: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
    leastereo_forward(sd, left, right, maxdisp, arch, training=True, batch_stats=stats)
    out = {k: v.clone() for k, v in sd.items()}
    for prefix, (mean, var_unbiased) in stats.items():
        out[prefix + ".bn.running_mean"] = mean
        out[prefix + ".bn.running_var"] = var_unbiased
    return out


def tolerance_report(d_test: torch.Tensor, d_ref: torch.Tensor) -> dict:
    """North-star tolerance: |dd| <= 0.1 px on >= 99.9 % of pixels and mean |dd| <= 0.01 px."""
    diff = (d_test.double() - d_ref.double()).abs()
    frac = float((diff <= 0.1).double().mean())
    return {"frac_within_0p1": frac, "mean_abs": float(diff.mean()), "max_abs": float(diff.max()),
            "ok": bool(frac >= 0.999 and float(diff.mean()) <= 0.01)}


# --------------------------------------------------------------------------------------------------------
# training step (train.py:130-160): train-mode BN, smooth-L1 on valid pixels, autograd gradients
# --------------------------------------------------------------------------------------------------------

def cost_volume_torch(x: torch.Tensor, y: torch.Tensor, maxdisp: int) -> torch.Tensor:
    """Differentiable statement of retrain/LEAStereo.py:34-48 (same slice assignments)."""
    B, C, H, W = x.shape
    D = int(maxdisp / 3)
    cost = x.new_zeros((B, 2 * C, D, H, W))
    for i in range(D):
        if i >= W:
            break
        if i > 0:
            cost[:, :C, i, :, i:] = x[:, :, :, i:]
            cost[:, C:, i, :, i:] = y[:, :, :, :-i]
        else:
            cost[:, :C, i] = x
            cost[:, C:, i] = y
    return cost


def hot_path_train(sd, fx: torch.Tensor, fy: torch.Tensor, maxdisp: int, arch=SHIPPED_ARCH, batch_stats=None):
    """Train-mode hot path on feature maps, differentiable w.r.t. fx, fy and every tensor of ``sd``."""
    cost = cost_volume_torch(fx, fy, maxdisp)
    mat = matching_forward(sd, cost, arch, training=True, batch_stats=batch_stats)
    return disp_head(mat, maxdisp), mat


def train_loss(disp: torch.Tensor, target: torch.Tensor, maxdisp: int) -> torch.Tensor:
    mask = (target < maxdisp) & (target > 0.001)                 # train.py:116-118
    return F.smooth_l1_loss(disp[mask], target[mask], reduction="mean")   # train.py:157
