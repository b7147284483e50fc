"""CPU tests (-m "not gpu") of the TRAIN-mode hot path (batch-statistics BN forward + full backward) through the CPU
emulation of the kernels, against the oracle's autograd gradients.  Exact storage (3 planes) + fp32 SIMT convs, so the
tolerances are tight."""
import pytest
import torch

import kernel_checks as K
from conftest import load_golden, golden_arch, golden_model, golden_state_dict, seeded_model
from oracle import leastereo_oracle as O
from test_emu_kernels import emu_ops  # noqa: F401  (fixture)

DEV = torch.device("cpu")


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def check_train_step(ops, device, planes=3, conv="simt", tol=2e-3, grad_tol=2e-2, name="cal_b2_24x48_d24"):
    g = load_golden(name)
    maxdisp = int(g["maxdisp"])
    model = golden_model(g, device).train()
    model.engine_options = {"planes": planes, "train_planes": planes, "conv": conv}
    fx = torch.from_numpy(g["fx"]).to(device).requires_grad_(True)
    fy = torch.from_numpy(g["fy"]).to(device).requires_grad_(True)
    gen = torch.Generator().manual_seed(3)
    target = (torch.rand(fx.shape[0], 24, 48, generator=gen) * maxdisp * 0.5).to(device)

    from leastereo_b200.training import hot_path_train_forward
    sd_before = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    disp = hot_path_train_forward(model, fx, fy, ops=ops)
    loss = O.train_loss(disp, target, maxdisp)
    loss.backward()

    # ---- oracle: same weights as leaves, autograd
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd_before.items()}
    ofx = torch.from_numpy(g["fx"]).requires_grad_(True)
    ofy = torch.from_numpy(g["fy"]).requires_grad_(True)
    stats = {}
    odisp, omat = O.hot_path_train(sd, ofx, ofy, maxdisp, golden_arch(g), batch_stats=stats)
    oloss = O.train_loss(odisp, target.cpu(), maxdisp)
    oloss.backward()

    # Gradients of this randomly initialised ReLU network are discontinuous in the activations (a rounding-level
    # change flips ReLU masks), so besides max-norm errors the check uses per-tensor relative L2 errors and the cosine
    # between the full gradient vectors - what an optimiser step actually sees.
    def l2rel(a, b):
        return float((a.double() - b.double()).norm()) / max(float(b.double().norm()), 1e-30)

    assert _rel(disp.detach().cpu(), odisp.detach()) <= tol, "train-mode forward"
    assert abs(float(loss.detach()) - float(oloss.detach())) <= tol * max(1.0, abs(float(oloss.detach())))
    stats_out = {"dfx_l2": l2rel(fx.grad.cpu(), ofx.grad), "dfy_l2": l2rel(fy.grad.cpu(), ofy.grad)}
    errs, dot, na, nb = [], 0.0, 0.0, 0.0
    worst = ("", 0.0)
    for name, p in model.matching.named_parameters():
        og = sd["matching." + name].grad
        if og is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, "missing gradient for " + name
        pg = p.grad.cpu().double()
        e = l2rel(pg, og)
        errs.append(e)
        if e > worst[1]:
            worst = (name, e)
        dot += float((pg * og.double()).sum()); na += float((pg * pg).sum()); nb += float((og.double() ** 2).sum())
    assert len(errs) > (250 if "matching_genotype" not in g else 200)
    errs.sort()
    stats_out.update(param_l2_median=errs[len(errs) // 2], param_l2_max=errs[-1], worst=worst[0],
                     cosine=dot / max((na * nb) ** 0.5, 1e-300))
    assert stats_out["dfx_l2"] <= grad_tol and stats_out["dfy_l2"] <= grad_tol, stats_out
    assert stats_out["param_l2_median"] <= grad_tol and stats_out["param_l2_max"] <= 10 * grad_tol, stats_out
    assert stats_out["cosine"] >= 1.0 - grad_tol, stats_out
    # running statistics follow momentum 0.1 with the unbiased batch variance
    for prefix in ("matching.stem0", "matching.cells.5._ops.3", "matching.cells.5._ops.2", "matching.last_6"):
        if prefix not in stats:                 # an Identity op of a skip_connect genotype has no BN
            continue
        mean, var_unbiased = stats[prefix]
        rm = 0.9 * sd_before[prefix + ".bn.running_mean"] + 0.1 * mean.detach()
        rv = 0.9 * sd_before[prefix + ".bn.running_var"] + 0.1 * var_unbiased.detach()
        assert _rel(model.state_dict()[prefix + ".bn.running_mean"].cpu(), rm) <= 1e-3
        assert _rel(model.state_dict()[prefix + ".bn.running_var"].cpu(), rv) <= 1e-3
        assert int(model.state_dict()[prefix + ".bn.num_batches_tracked"]) == int(sd_before[prefix + ".bn.num_batches_tracked"]) + 1
    return stats_out


def test_train_step_matches_autograd(emu_ops):
    print("train step vs autograd:", check_train_step(emu_ops, DEV))


def test_train_step_skip_connect_genotype(emu_ops):
    """3D genotype with skip_connect ops (operations_3d.py:84-104, genotypes_3d.py:5-8): Identity inside the step sums,
    forward and backward, against the oracle's autograd (the oracle itself is pinned to the reference's run of this
    genotype by tests/golden/cal_skip_b2_24x48_d24.npz)."""
    print("skip genotype train step vs autograd:", check_train_step(emu_ops, DEV, name="cal_skip_b2_24x48_d24"))


def test_backward_kernels_individually(emu_ops):
    check_backward_kernels(emu_ops, DEV)


def check_backward_kernels(ops, DEV):
    """Trilinear backward, cost-volume backward and disparity-head backward against autograd on random data (run on the
    CPU emulation here and on the GPU by tests/test_gpu_training.py)."""
    import torch.nn.functional as F
    from leastereo_b200.kernels import PlanesVol
    for src_sp, dst_sp in [((4, 3, 6), (8, 6, 12)), ((8, 6, 12), (4, 3, 6)), ((3, 5, 4), (5, 9, 7)), ((2, 2, 2), (1, 1, 1))]:
        x = torch.randn(1, 8, *src_sp, requires_grad=True)
        y = F.interpolate(x, dst_sp, mode="trilinear", align_corners=True)
        gy = torch.randn_like(y)
        y.backward(gy)
        ddst = ops.pack(gy.to(DEV), 3)
        dsrc = PlanesVol.empty(1, 8, 3, *src_sp, DEV)
        dsrc.t.zero_()
        ops.trilinear_ac_bwd(ddst, 0, dsrc, 0, 8)
        assert _rel(ops.unpack(dsrc).cpu(), x.grad) <= 1e-5, (src_sp, dst_sp)
    fx = torch.randn(2, 8, 5, 12, requires_grad=True)
    fy = torch.randn(2, 8, 5, 12, requires_grad=True)
    cost = O.cost_volume_torch(fx, fy, 15)
    gc = torch.randn_like(cost)
    cost.backward(gc)
    dx, dy = ops.cost_volume_bwd(ops.pack(gc.to(DEV), 3), 8)
    assert _rel(dx.cpu(), fx.grad) <= 1e-5 and _rel(dy.cpu(), fy.grad) <= 1e-5
    # disparity head backward
    mat = (torch.randn(1, 1, 8, 5, 7) * 2).requires_grad_(True)
    for maxdisp in (24, 25):
        mat.grad = None
        d = O.disp_head(mat, maxdisp)
        go = torch.randn_like(d)
        d.backward(go)
        dm = ops.disp_head_bwd(mat.detach().to(DEV), go.to(DEV), maxdisp)
        assert _rel(dm.cpu(), mat.grad[:, 0]) <= 2e-4, maxdisp


def test_gradient_coverage_analysis():
    """training._Coverage: the first writer of a gradient slice overwrites, a later writer of covered channels
    accumulates, partial overlaps and reads of incompletely written slices fall back to a zero-filled volume."""
    from leastereo_b200.training import _Coverage
    from leastereo_b200.engine import Slice

    class Vol:
        def __init__(self, c):
            self.C = c

    v, w = Vol(48), Vol(32)
    cov = _Coverage(set())
    assert cov.write(Slice(v, 0, 16)) is False            # first writer of [0, 16)
    assert cov.write(Slice(v, 16, 16)) is False           # disjoint range: also a first writer
    assert cov.write(Slice(v, 0, 32)) is True             # fully covered by the two ranges: accumulates
    assert not cov.retry
    cov.read(Slice(v, 0, 32))                             # completely written: fine
    assert not cov.retry
    cov.read(Slice(v, 0, 48))                             # [32, 48) never written: the volume must be zero-filled
    assert cov.retry and id(v) in cov.zero_vols
    cov2 = _Coverage(cov.zero_vols)
    assert cov2.write(Slice(v, 0, 16)) is True            # zero-filled volumes: every writer accumulates
    assert cov2.write(Slice(w, 0, 16)) is False
    assert cov2.write(Slice(w, 8, 16)) is True            # partial overlap [8, 24) over [0, 16): fall back, retry
    assert cov2.retry and id(w) in cov2.zero_vols


def test_train_plan_zero_fills_one_volume_for_the_shipped_genotype(emu_ops):
    """With the reference's architecture every gradient volume but one has first writers covering it.  The exception is
    the shared [C1 | C4 | C8] skip buffer: conv2's data gradient writes [C4 | C8] first and conv1's then writes
    [C1 | C4], a partial overlap, so that one volume is zero-filled and its writers accumulate."""
    from leastereo_b200.training import TrainPlan
    model = seeded_model(24)
    plan = TrainPlan(model.matching, emu_ops, 1, (8, 8, 16), 3, DEV, "simt", 24)
    assert len(plan.zero_grads) == 1 and plan.zero_grads[0].C == 3 * 64
    assert len(plan.values) > 120
    assert sum(1 for n in plan.nodes if getattr(n, "up", False)) >= 6      # the up-sampling cells take conv-before-upsample
