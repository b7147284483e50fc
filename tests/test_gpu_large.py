"""GPU parity tests (-m gpu) at BASELINE.json sizes against outputs of the UNMODIFIED reference (fp32, CPU) stored in
tests/golden/large_*.npz (written by tests/golden/make_golden_large.py in the build container; nothing here reads
/root/reference).  The engine runs with its DEFAULT options - tcgen05 convs on 2 bf16 planes (bf16x3), collapsed
stem0, fused cost volume, head-tap projection, native feature net - i.e. exactly what bench.py times.

Tolerance = BASELINE.json north_star: per-pixel |dd| <= 0.1 px on >= 99.9 % of pixels and mean |dd| <= 0.01 px.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden, seeded_model
from oracle import leastereo_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as g
    g.build()
    from leastereo_b200.kernels import get_ops
    return get_ops()


def pair_inputs(k, B, H, W):
    """Pair k of tests/golden/make_golden_large.py (seeded randn, left then right)."""
    g = torch.Generator().manual_seed(1 + k)
    return torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)


def sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def large_case(name):
    g = load_golden(name)
    H, W, maxdisp = int(g["H"]), int(g["W"]), int(g["maxdisp"])
    l0, r0 = pair_inputs(0, 1, H, W)
    l1, r1 = pair_inputs(1, 1, H, W)
    assert [sha(l0), sha(r0), sha(l1), sha(r1)] == [str(s) for s in g["input_sha256"]], "seeded inputs differ"
    model = seeded_model(maxdisp)
    model.load_state_dict(golden_state_dict(g, model))
    return g, model.to(DEV).eval(), (l0, r0, l1, r1)


def report(got, want):
    rep = O.tolerance_report(got.cpu(), want)
    return rep


@pytest.mark.parametrize("name", ["large_cal_288x576_d192", "large_raw_288x576_d192", "large_cal_384x1248_d192"])
def test_default_engine_batch1_vs_reference(ops, name):
    """configs[0] (288x576) in both parity regimes and KITTI 384x1248 calibrated, one pair per call, eager launches."""
    g, model, (l0, r0, l1, r1) = large_case(name)
    from leastereo_b200 import engine
    with torch.no_grad():
        d0 = model(l0.to(DEV), r0.to(DEV))
        plan = next(p for p in engine._plans(model.matching).values())
        assert plan.conv_mode == "tc" and plan.P == 2 and plan.fxy3 is not None, "not the default (collapsed, tc, P=2) plan"
        assert any(isinstance(p, engine.FeaturePlan) for p in engine._plans(model.feature).values())
        mat = plan.mat.cpu().numpy()
        d1 = model(l1.to(DEV), r1.to(DEV))
    rep0, rep1 = report(d0, torch.from_numpy(g["disp0"])), report(d1, torch.from_numpy(g["disp1"]))
    sample = mat[:, :, ::3, ::5, ::7]
    rep0["mat_rel_err"] = float(np.abs(sample - g["mat_sample"]).max() / float(g["mat_absmax"]))
    print(name, "pair0", rep0, "pair1", rep1)
    assert rep0["ok"], rep0
    assert rep1["ok"], rep1


@pytest.mark.parametrize("name", ["large_cal_288x576_d192", "large_cal_384x1248_d192"])
def test_default_engine_batch4_cuda_graph_vs_reference(ops, name):
    """Four pairs per step replayed from a CUDA graph (what bench.py times): pairs [0, 1, 0, 1]."""
    g, model, (l0, r0, l1, r1) = large_case(name)
    model.engine_options = {"assume_frozen": True}
    left = torch.cat([l0, l1, l0, l1]).to(DEV)
    right = torch.cat([r0, r1, r0, r1]).to(DEV)
    want = torch.cat([torch.from_numpy(g["disp0"]), torch.from_numpy(g["disp1"])] * 2)
    with torch.no_grad():
        model(left, right)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            model(left, right)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = model(left, right)
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
    rep = report(out, want)
    print(name, "B=4 graph", rep)
    assert rep["ok"], rep
    assert torch.equal(out[0], out[2]) and torch.equal(out[1], out[3]), "a pair's result depends on its batch slot"


def test_train_step_288x576_b4_vs_reference(ops):
    """configs[4]: fwd + bwd of one train.py:153-158 step at 288x576, batch 4, seeded init, ``model.train()``, against
    the reference's autograd run (loss, train-mode disparity sample, the gradient of EVERY parameter by norm / sum,
    selected gradients element-wise, updated BN running statistics)."""
    g = load_golden("large_train_288x576_b4")
    H, W, B, maxdisp = int(g["H"]), int(g["W"]), int(g["B"]), int(g["maxdisp"])
    gen = torch.Generator().manual_seed(1)
    left = torch.randn(B, 3, H, W, generator=gen)
    right = torch.randn(B, 3, H, W, generator=gen)
    target = torch.rand(B, H, W, generator=gen) * maxdisp * 0.5
    assert [sha(left), sha(right), sha(target)] == [str(s) for s in g["input_sha256"]]
    model = seeded_model(maxdisp)
    from conftest import state_dict_sha256
    assert state_dict_sha256(model.state_dict()) == str(g["state_sha256_init"])
    model = model.to(DEV).train()
    from leastereo_b200.pipeline import masked_smooth_l1_loss
    left_d = left.to(DEV).requires_grad_(True)
    right_d = right.to(DEV).requires_grad_(True)
    disp = model(left_d, right_d)
    loss = masked_smooth_l1_loss(disp, target.to(DEV), maxdisp)
    loss.backward()
    torch.cuda.synchronize()

    def l2rel(a, b):
        a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
        return float((a - b).norm()) / max(float(b.norm()), 1e-30)

    out = {"loss": float(loss), "ref_loss": float(g["loss"])}
    out["disp_l2"] = l2rel(disp.detach().cpu()[:, ::4, ::4], g["disp_sample"])
    out["dleft_l2"] = l2rel(left_d.grad.cpu()[:, :, ::8, ::8], g["dleft_sample"])
    names = [str(n) for n in g["param_names"]]
    stats = g["grad_stats"]
    params = dict(model.named_parameters())
    norm_err, dot_n, worst = [], 0, ("", 0.0)
    for n, (ref_norm, ref_sum) in zip(names, stats):
        p = params[n]
        if np.isnan(ref_norm):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        assert p.grad is not None, n
        e = abs(float(p.grad.double().norm()) - ref_norm) / max(ref_norm, 1e-30)
        norm_err.append(e)
        if e > worst[1]:
            worst = (n, e)
    norm_err.sort()
    out["grad_norm_err_median"], out["grad_norm_err_max"], out["worst"] = norm_err[len(norm_err) // 2], norm_err[-1], worst[0]
    full = {}
    for k in g:
        if k.startswith("grad/"):
            full[k[5:]] = l2rel(params[k[5:]].grad.cpu(), g[k])
    out["full_grad_l2_max"] = max(full.values())
    out["full_grad_l2"] = {k: round(v, 5) for k, v in full.items()}
    print("train step 288x576 B=4 vs reference:", out)
    # tolerances: the tensor-core path computes bf16x3 products on 2-plane storage; gradients of this random-init ReLU
    # net flip masks on rounding-level changes, hence relative L2 / norm checks (as tests/test_training_emu.py)
    assert abs(out["loss"] - out["ref_loss"]) <= 2e-3 * abs(out["ref_loss"]), out
    assert out["disp_l2"] <= 5e-3, out
    assert out["dleft_l2"] <= 5e-2, out
    assert out["grad_norm_err_median"] <= 1e-2 and out["grad_norm_err_max"] <= 1e-1, out
    assert out["full_grad_l2_max"] <= 5e-2, out
    sd = model.state_dict()
    for k in g:
        if k.startswith("bn_after/"):
            got, want = sd[k[9:]].cpu(), torch.from_numpy(g[k])
            assert float((got - want).abs().max()) <= 2e-3 * max(1.0, float(want.abs().max())), k
