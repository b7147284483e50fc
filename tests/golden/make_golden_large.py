"""Golden vectors at BASELINE.json sizes, produced by running the UNMODIFIED reference (/root/reference) on CPU (fp32)
in the build container:

    python tests/golden/make_golden_large.py [case ...]      # writes tests/golden/large_*.npz

Inference cases (configs[0] 288x576 and configs[2] KITTI 384x1248, maxdisp 192).  Inputs and weights are NOT stored:
pair k is ``randn`` from ``torch.Generator().manual_seed(1 + k)`` (left then right, as make_golden.make_inputs) and the
weights are the ``torch.manual_seed(0)`` init; both are pinned by SHA-256.  Stored: the reference's disparity for
pairs 0 and 1 (pair 1 runs through the statistics calibrated on pair 0, so a batch can mix them), a strided sample +
SHA-256 of pair 0's matching volume ``mat``, and (calibrated regime) every BN running statistic.

Training case (configs[4]: 288x576, batch 4, ``model.train()``, seeded init): one ``train.py:153-158`` step - loss,
a strided sample of the train-mode disparity, per-parameter gradient statistics (L2 norm and sum) for EVERY parameter,
full gradients of a few named parameters, and the updated running statistics of a few BN layers.
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference, state_dict_sha256  # noqa: E402

INFER_CASES = [
    # name, H, W, maxdisp, regime
    ("large_cal_288x576_d192", 288, 576, 192, "calibrated"),
    ("large_raw_288x576_d192", 288, 576, 192, "raw"),
    ("large_cal_384x1248_d192", 384, 1248, 192, "calibrated"),
]
TRAIN_CASE = ("large_train_288x576_b4", 288, 576, 192, 4)
FULL_GRADS = ["matching.stem0.conv.weight", "matching.stem1.conv.weight", "matching.cells.0._ops.0.conv.weight",
              "matching.cells.5._ops.3.conv.weight", "matching.cells.10._ops.1.conv.weight",
              "matching.cells.6.preprocess.conv.weight", "matching.conv1.conv.weight", "matching.last_6.conv.weight",
              "matching.last_3.conv.weight", "matching.stem0.bn.weight", "matching.stem0.bn.bias",
              "matching.cells.3._ops.2.bn.weight", "feature.stem0.conv.weight", "feature.last_3.conv.weight"]
BN_AFTER = ["matching.stem0", "matching.cells.5._ops.3", "matching.last_6", "feature.stem1"]


def pair_inputs(k, B, H, W):
    g = torch.Generator().manual_seed(1 + k)
    return torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)


def sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def infer_case(name, H, W, maxdisp, regime):
    t0 = time.time()
    model = build_reference(maxdisp)
    out = dict(maxdisp=np.int64(maxdisp), regime=np.array(regime), H=np.int64(H), W=np.int64(W),
               state_sha256_init=np.array(state_dict_sha256(model.state_dict())),
               torch_version=np.array(torch.__version__))
    l0, r0 = pair_inputs(0, 1, H, W)
    l1, r1 = pair_inputs(1, 1, H, W)
    out["input_sha256"] = np.array([sha(l0), sha(r0), sha(l1), sha(r1)])
    if regime == "calibrated":
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                m.momentum = 1.0
        model.train()
        with torch.no_grad():
            model(l0, r0)
        for k, v in model.state_dict().items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                out["bn/" + k] = v.numpy().copy()
    model.eval()
    stages = {}
    h = model.matching.register_forward_hook(lambda m, i, o: stages.update(mat=o))
    with torch.no_grad():
        d0 = model(l0, r0)
        mat = stages["mat"].numpy().copy()
        d1 = model(l1, r1)
    h.remove()
    out["disp0"], out["disp1"] = d0.numpy(), d1.numpy()
    out["mat_sha256"] = np.array(hashlib.sha256(mat.tobytes()).hexdigest())
    out["mat_shape"] = np.array(mat.shape)
    out["mat_sample"] = mat[:, :, ::3, ::5, ::7].copy()
    out["mat_absmax"] = np.float64(np.abs(mat).max())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "disp0 mean %.4f std %.4f" % (d0.mean(), d0.std()), "mat std %.4g" % mat.std(),
          os.path.getsize(path) // 1024, "KiB", "%.0f s" % (time.time() - t0), flush=True)


def train_case(name, H, W, maxdisp, B):
    t0 = time.time()
    import torch.nn.functional as F
    model = build_reference(maxdisp)
    out = dict(maxdisp=np.int64(maxdisp), H=np.int64(H), W=np.int64(W), B=np.int64(B),
               state_sha256_init=np.array(state_dict_sha256(model.state_dict())),
               torch_version=np.array(torch.__version__))
    g = torch.Generator().manual_seed(1)
    left = torch.randn(B, 3, H, W, generator=g).requires_grad_(True)       # train.py:137-138
    right = torch.randn(B, 3, H, W, generator=g).requires_grad_(True)
    target = torch.rand(B, H, W, generator=g) * maxdisp * 0.5
    out["input_sha256"] = np.array([sha(left), sha(right), sha(target)])
    mask = (target < maxdisp) & (target > 0.001)                           # train.py:116-118
    model.train()
    disp = model(left, right)
    loss = F.smooth_l1_loss(disp[mask], target[mask], reduction="mean")    # train.py:157
    loss.backward()
    out["loss"] = np.float64(loss.item())
    out["error"] = np.float64(torch.mean(torch.abs(disp[mask] - target[mask])).item())     # train.py:162
    out["disp_sample"] = disp.detach()[:, ::4, ::4].numpy().copy()
    out["dleft_sample"] = left.grad[:, :, ::8, ::8].numpy().copy()
    out["dleft_norm"] = np.float64(left.grad.double().norm().item())
    out["dright_norm"] = np.float64(right.grad.double().norm().item())
    names, stats = [], []
    for k, p in model.named_parameters():
        names.append(k)
        if p.grad is None:
            stats.append([np.nan, np.nan])
        else:
            stats.append([p.grad.double().norm().item(), p.grad.double().sum().item()])
    out["param_names"] = np.array(names)
    out["grad_stats"] = np.array(stats, dtype=np.float64)
    sdp = dict(model.named_parameters())
    for k in FULL_GRADS:
        out["grad/" + k] = sdp[k].grad.numpy().copy()
    sd = model.state_dict()
    for prefix in BN_AFTER:
        out["bn_after/" + prefix + ".bn.running_mean"] = sd[prefix + ".bn.running_mean"].numpy().copy()
        out["bn_after/" + prefix + ".bn.running_var"] = sd[prefix + ".bn.running_var"].numpy().copy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss %.6f error %.6f" % (out["loss"], out["error"]), os.path.getsize(path) // 1024, "KiB",
          "%.0f s" % (time.time() - t0), flush=True)


def main():
    torch.set_num_threads(int(os.environ.get("LEA_GOLDEN_THREADS", "8")))
    want = set(sys.argv[1:])
    for case in INFER_CASES:
        if not want or case[0] in want:
            infer_case(*case)
    if not want or TRAIN_CASE[0] in want:
        train_case(*TRAIN_CASE)


if __name__ == "__main__":
    main()
