"""GPU parity tests (-m gpu): every kernel through the real C-ABI library (libleastereo_b200.so) against the oracle
and the golden vectors.  Nothing here reads /root/reference."""
import os

import pytest
import torch

import kernel_checks as K
from conftest import load_golden
from oracle import leastereo_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as g
    g.build()
    from leastereo_b200.kernels import get_ops
    return get_ops()


DEV = torch.device("cuda:0")


def test_device_build(ops):
    assert ops.device_build and not ops.missing


def test_cost_volume_f32_bit_exact(ops):
    K.check_cost_volume_f32(ops, DEV)


@pytest.mark.parametrize("name", ["cal_48x96_d48", "cal_46x94_d50", "cal_b2_24x48_d24"])
def test_cost_volume_golden(ops, name):
    K.check_cost_volume_golden(ops, DEV, name)


def test_pack_unpack(ops):
    K.check_pack_unpack(ops, DEV)


def test_cost_volume_planes(ops):
    K.check_cost_volume_planes(ops, DEV)


def test_trilinear(ops):
    K.check_trilinear(ops, DEV)


def test_resample_conv1x1(ops):
    K.check_resample_conv1x1(ops, DEV)


def test_conv_simt(ops):
    K.check_conv_simt(ops, DEV)


def test_disp_head(ops):
    K.check_disp_head(ops, DEV)


def test_head_taps(ops):
    K.check_head_taps(ops, DEV)


def test_stem0_collapse(ops):
    print("collapsed stem0 vs conv3d(cost volume): rel err", K.check_stem0_collapse(ops, DEV))


def test_disparity_regression(ops):
    K.check_disparity_regression(ops, DEV)


@pytest.mark.parametrize("name", ["raw_48x96_d48", "cal_48x96_d48", "cal_46x94_d50", "cal_b2_24x48_d24",
                                  "cal_skip_b2_24x48_d24"])
@pytest.mark.parametrize("planes", [3, 2])
def test_hot_path_golden_simt(ops, name, planes):
    rep = K.check_hot_path_golden(ops, DEV, name, planes=planes, conv="simt", mat_rtol=None)
    print(name, planes, rep)


# ---- tcgen05 path ------------------------------------------------------------------------------------------
def test_tc_selftest(ops):
    rc = ops.tc_selftest(1)
    assert rc == 0, ops.lib.lea_last_error().decode()


@pytest.mark.parametrize("planes", [2, 3, 1])
def test_conv_tc(ops, planes):
    K.check_conv_tc(ops, DEV, planes=planes, verbose=True)


@pytest.mark.parametrize("planes", [2, 3])
def test_conv_tc_depth_chunks(ops, planes):
    """Forced depth-chunk lengths, ragged tiles, 16+8 channel groups and wide 1x1x1 tiles."""
    worst = K.check_conv_tc(ops, DEV, planes=planes, cases=K.TC_CHUNK_CASES, verbose=True)
    print("worst relative error, planes", planes, worst)


def test_conv_tc_single_pass(ops):
    K.check_conv_tc(ops, DEV, planes=2, mma_terms=1, verbose=True)


@pytest.mark.parametrize("name", ["raw_48x96_d48", "cal_48x96_d48", "cal_46x94_d50", "cal_b2_24x48_d24",
                                  "cal_skip_b2_24x48_d24"])
@pytest.mark.parametrize("planes", [2, 3])
def test_hot_path_golden_tc(ops, name, planes):
    rep = K.check_hot_path_golden(ops, DEV, name, planes=planes, conv="tc", mat_rtol=None)
    print(name, planes, rep)


def test_fused_cost_volume_equals_materialised(ops):
    """stem0 with the cost volume built inside its TMA loader must reproduce the run on the materialised volume
    bit for bit (same operands, same MMA order)."""
    g = load_golden("cal_48x96_d48")
    outs = {}
    for fuse_cv in (False, True):
        model = K.seeded_model(int(g["maxdisp"]))
        model.load_state_dict(K.golden_state_dict(g, model))
        model = model.to(DEV).eval()
        model.engine_options = {"planes": 2, "conv": "tc", "fuse_cv": fuse_cv, "collapse_stem0": False}
        fx, fy = torch.from_numpy(g["fx"]).to(DEV), torch.from_numpy(g["fy"]).to(DEV)
        from leastereo_b200 import engine
        disp = engine.hot_path_forward(model, fx, fy, ops=ops)
        plan = engine.get_plan(model.matching, 1, (16, 16, 32), fx.device, engine._options(model), ops)
        assert (plan.fxp is not None) == fuse_cv
        outs[fuse_cv] = (plan.mat.clone(), disp.clone())
    assert torch.equal(outs[False][0], outs[True][0])
    assert torch.equal(outs[False][1], outs[True][1])


# ---- BASELINE.json full sizes: size-independent properties -----------------------------------------------------
def _random_model(maxdisp, options):
    model = K.seeded_model(maxdisp).to(DEV).eval()
    model.engine_options = dict(options)
    return model


def test_stem0_collapse_tc_covers_every_voxel(ops):
    print("assemble + band/edge tensor-core launch vs conv3d(cost volume): rel err", K.check_stem0_collapse_tc(ops, DEV))


def test_collapsed_stem0_matches_plain(ops):
    """stem0 through the 2-D maps + assemble + band/edge tensor-core launch must agree with the plain fused-loader run
    on every voxel (different summation order: compare the stem0 output volume to fp32 rounding of 2 planes)."""
    from leastereo_b200 import engine
    g = load_golden("cal_48x96_d48")
    vols, disps = {}, {}
    for collapse in (False, True):
        model = K.seeded_model(int(g["maxdisp"]))
        model.load_state_dict(K.golden_state_dict(g, model))
        model = model.to(DEV).eval()
        model.engine_options = {"planes": 2, "conv": "tc", "collapse_stem0": collapse,
                                "reuse_buffers": False}      # the stem0 volume is inspected after the run
        fx, fy = torch.from_numpy(g["fx"]).to(DEV), torch.from_numpy(g["fy"]).to(DEV)
        disps[collapse] = engine.hot_path_forward(model, fx, fy, ops=ops).cpu()
        plan = engine.get_plan(model.matching, 1, (16, 16, 32), fx.device, engine._options(model), ops)
        assert (plan.fxy3 is not None) == collapse
        vols[collapse] = ops.unpack(plan.volumes[0]).cpu()            # v0 = stem0's output
    err = float((vols[True] - vols[False]).abs().max()) / float(vols[False].abs().max())
    print("collapsed vs plain stem0: rel err", err)
    assert err <= 1e-4
    assert O.tolerance_report(disps[True], disps[False])["ok"]


@pytest.mark.parametrize("conv", ["tc", "simt"])
def test_fused_resample_conv_is_on_the_plan_and_matches(ops, conv):
    """lea_resample_conv1x1 on the plan (default) against the resample + conv launches it replaces."""
    g = load_golden("cal_48x96_d48")
    outs = {}
    for fused in (False, True):
        mat, disp, model, plan = K.run_hot_path(ops, DEV, g, planes=3, conv=conv, extra={"fuse_resample_conv": fused})
        kinds = [s.kind for s in plan.steps]
        assert ("resample_conv1x1" in kinds) == fused
        outs[fused] = (mat, disp)
    assert float((outs[True][1] - outs[False][1]).abs().max()) <= 2e-3
    assert float((outs[True][0] - outs[False][0]).abs().max()) <= 5e-4 * float(outs[False][0].abs().max())


def test_cost_volume_full_kitti_bit_exact(ops):
    """KITTI 1/3-res shape (32 x 128 x 416, D3 = 64): 886 MB volume, bit-exact against the oracle's numpy loop."""
    g = torch.Generator().manual_seed(11)
    x, y = torch.randn(1, 32, 128, 416, generator=g), torch.randn(1, 32, 128, 416, generator=g)
    got = ops.cost_volume_f32(x.to(DEV), y.to(DEV), 192).cpu().numpy()
    want = O.cost_volume_numpy(x.numpy(), y.numpy(), 192)
    assert got.tobytes() == want.tobytes()
    # the planes-layout kernel must hold the same values (3 planes = exact fp32)
    vol = ops.cost_volume_planes(x.to(DEV), y.to(DEV), 192, 3)
    assert torch.equal(ops.unpack(vol).cpu(), torch.from_numpy(want))


def test_disp_head_full_kitti(ops):
    g = torch.Generator().manual_seed(12)
    mat = torch.randn(1, 1, 64, 128, 416, generator=g) * 4
    got = ops.disp_head(mat.to(DEV), 192).cpu()
    want = O.disp_head(mat, 192)
    assert float((got - want).abs().max()) <= 2e-3


def test_hot_path_kitti_tc_vs_fp32_simt(ops):
    """Full KITTI 384x1248, D=192, random-init weights: the tensor-core bf16x3 path against the fp32 SIMT path with
    exact (3-plane) storage - the mode that is pinned to the reference by the golden tests."""
    torch.manual_seed(5)
    left, right = torch.randn(1, 3, 384, 1248, device=DEV), torch.randn(1, 3, 384, 1248, device=DEV)
    outs = {}
    for name, opt in (("tc", {"planes": 2, "conv": "tc"}), ("simt", {"planes": 3, "conv": "simt"})):
        model = _random_model(192, opt)
        with torch.no_grad():
            outs[name] = model(left, right).float().cpu()
        del model
        torch.cuda.empty_cache()
    rep = O.tolerance_report(outs["tc"], outs["simt"])
    print("KITTI tc(P=2) vs simt(P=3):", rep)
    assert outs["tc"].shape == (1, 384, 1248) and torch.isfinite(outs["tc"]).all()
    assert rep["ok"], rep


def test_batch_independence_sceneflow(ops):
    """SceneFlow 576x960 (configs[1]): in eval mode a pair's disparity must not depend on its batch-mates."""
    torch.manual_seed(6)
    left, right = torch.randn(3, 3, 576, 960, device=DEV), torch.randn(3, 3, 576, 960, device=DEV)
    model = _random_model(192, {"planes": 2, "conv": "tc"})
    with torch.no_grad():
        full = model(left, right)
        one = model(left[1:2], right[1:2])
    assert full.shape == (3, 576, 960)
    assert torch.equal(full[1:2], one)


def test_middlebury_shape_runs(ops):
    """Middlebury half-res 1008x1512, maxdisp 408 (configs[3], the largest cost volume): runs, finite, in range."""
    torch.manual_seed(7)
    left, right = torch.randn(1, 3, 1008, 1512, device=DEV), torch.randn(1, 3, 1008, 1512, device=DEV)
    model = _random_model(408, {"planes": 2, "conv": "tc"})
    with torch.no_grad():
        d = model(left, right)
    assert d.shape == (1, 1008, 1512) and torch.isfinite(d).all()
    assert float(d.min()) >= 0.0 and float(d.max()) <= 407.0


def test_ragged_input_size(ops):
    """H, W not multiples of 3 and maxdisp % 3 != 0 (output is 3*ceil(H/3) x 3*ceil(W/3) like the reference)."""
    rep = K.check_hot_path_golden(ops, DEV, "cal_46x94_d50", planes=2, conv="tc", mat_rtol=None)
    g = load_golden("cal_46x94_d50")
    model = K.seeded_model(50)
    model.load_state_dict(K.golden_state_dict(g, model))
    model = model.to(DEV).eval()
    with torch.no_grad():
        d = model(torch.from_numpy(g["left"]).to(DEV), torch.from_numpy(g["right"]).to(DEV))
    assert d.shape == (1, 48, 96)
    full = O.tolerance_report(d.cpu(), torch.from_numpy(g["disp"]))      # includes the stock-PyTorch feature net
    assert full["ok"], full


def test_state_dict_roundtrip_and_module_prefix(ops):
    g = load_golden("cal_b2_24x48_d24")
    model = K.seeded_model(24)
    sd = K.golden_state_dict(g, model)
    model.load_state_dict({"module." + k: v for k, v in sd.items()})      # DataParallel-style checkpoint (train.py:376)
    model = model.to(DEV).eval()
    with torch.no_grad():
        d = model(torch.from_numpy(g["left"]).to(DEV), torch.from_numpy(g["right"]).to(DEV))
    assert O.tolerance_report(d.cpu(), torch.from_numpy(g["disp"]))["ok"]
    # parameters changed in place must be picked up (BN refresh + weight re-pack)
    with torch.no_grad():
        model.matching.stem0.conv.weight.mul_(1.5)
        d2 = model(torch.from_numpy(g["left"]).to(DEV), torch.from_numpy(g["right"]).to(DEV))
    assert not torch.equal(d, d2)


# ---- native feature net ------------------------------------------------------------------------
def test_native_feature_net_simt(ops):
    assert K.check_feature_plan(ops, DEV, "cal_46x94_d50", planes=3, conv="simt") <= 2e-4


def test_native_feature_net_tc(ops):
    err = K.check_feature_plan(ops, DEV, "cal_48x96_d48", planes=2, conv="tc", tol=5e-4)
    print("native feature net (tc, 2 planes) rel err", err)


@pytest.mark.parametrize("name", ["raw_48x96_d48", "cal_48x96_d48", "cal_46x94_d50", "cal_b2_24x48_d24"])
def test_full_forward_native_vs_golden(ops, name):
    """Images in, disparity out, everything on the native kernels (default engine options) against the reference."""
    g = load_golden(name)
    model = K.seeded_model(int(g["maxdisp"]))
    model.load_state_dict(K.golden_state_dict(g, model))
    model = model.to(DEV).eval()
    with torch.no_grad():
        d = model(torch.from_numpy(g["left"]).to(DEV), torch.from_numpy(g["right"]).to(DEV))
    from leastereo_b200 import engine
    assert any(isinstance(p, engine.FeaturePlan) for p in engine._plans(model.feature).values()), "native feature path not taken"
    rep = O.tolerance_report(d.cpu(), torch.from_numpy(g["disp"]))
    print(name, rep)
    assert rep["ok"], rep


def test_pack_weights_dgrad_equals_explicit_transpose(ops):
    """lea_pack_weights_tc_dgrad (the data-gradient conv's weight image, read in place from the forward weight) is
    bit-identical to packing the explicitly flipped + transposed weight, for whole tensors and 64-channel slices."""
    g = torch.Generator().manual_seed(5)
    for (co, ci, k) in [(32, 64, 3), (64, 128, 3), (16, 16, 3), (8, 8, 3), (16, 64, 1), (32, 128, 1), (8, 64, 1)]:
        w = torch.randn(co, ci, k, k, k, generator=g).to(DEV)
        wt = w.flip(2, 3, 4).transpose(0, 1).contiguous()               # (ci, co, k, k, k): the data-gradient conv's weight
        for o0 in range(0, ci, 64):
            oc = min(64, ci - o0)
            if ops.tc_weight_image_bytes(co, oc, k, 2) <= 0:
                continue
            want = ops.pack_weights_tc(wt[o0:o0 + oc].contiguous(), 2)
            got = ops.pack_weights_tc_dgrad(w, 2, o0, oc)
            assert torch.equal(got, want), (co, ci, k, o0)


def test_module_owned_cuda_graph(ops):
    """LEAStereo.forward replays its own CUDA graph from the third call on (engine option cuda_graph, default on): same
    bits as the eager launch list, fresh inputs honoured, in-place parameter changes picked up, no aliasing of results."""
    g = load_golden("cal_48x96_d48")
    model = K.seeded_model(int(g["maxdisp"]))
    model.load_state_dict(K.golden_state_dict(g, model))
    model = model.to(DEV).eval()
    left, right = torch.from_numpy(g["left"]).to(DEV), torch.from_numpy(g["right"]).to(DEV)
    from leastereo_b200 import engine
    with torch.no_grad():
        model.engine_options = {"cuda_graph": False}
        eager = model(left, right).clone()
        eager_swapped = model(right, left).clone()
        model.engine_options = {}
        outs = [model(left, right) for _ in range(5)]
        fplan = [p for p in engine._plans(model.feature).values() if isinstance(p, engine.FeaturePlan)][-1]
        assert fplan.__dict__["_graph"]["graph"] is not None, "the module did not capture its graph"
        for o in outs:
            assert torch.equal(o, eager)
        swapped = model(right, left)                       # new input values through the static buffers
        assert torch.equal(swapped, eager_swapped) and torch.equal(outs[-1], eager)     # earlier results are not aliased
        model.matching.stem1.conv.weight.mul_(1.25)         # parameters changed in place: re-packed, graph re-captured
        changed = model(left, right)
        model.engine_options = {"cuda_graph": False}
        want = model(left, right)
    assert not torch.equal(changed, eager) and torch.equal(changed, want)


# ---- callers on either side of the path ------------------------------------------------
def test_normalize_pad(ops):
    K.check_normalize_pad(ops, DEV)


def test_masked_smooth_l1(ops):
    K.check_masked_smooth_l1(ops, DEV)
    # the autograd wrapper used by a training script
    from leastereo_b200.pipeline import masked_smooth_l1_loss
    from oracle import io_oracle as IO
    disp, target, maxdisp = K._loss_case(DEV, seed=9)
    d = disp.clone().requires_grad_(True)
    loss = masked_smooth_l1_loss(d, target, maxdisp)
    (loss * 3.0).backward()
    r = disp.detach().cpu().clone().requires_grad_(True)
    want = IO.masked_smooth_l1(r, target.cpu(), maxdisp)
    (want * 3.0).backward()
    assert abs(float(loss) - float(want)) <= 1e-6 and float((d.grad.cpu() - r.grad).abs().max()) <= 1e-7


def test_flat_adam(ops):
    K.check_adam(ops, DEV)


def test_disparity_metrics(ops):
    K.check_disparity_metrics(ops, DEV)


def test_input_pipeline_double_buffer(ops):
    """InputPipeline: pinned uint8 uploads on a side stream, normalised on the compute stream; KITTI-sized pairs."""
    import numpy as np
    from oracle import io_oracle as IO
    from leastereo_b200.pipeline import InputPipeline
    g = np.random.RandomState(1)
    H, W, ch, cw = 75, 124, 96, 144
    pairs = [(g.randint(0, 256, (H, W, 3)).astype(np.uint8), g.randint(0, 256, (H, W, 3)).astype(np.uint8)) for _ in range(5)]
    pipe = InputPipeline(H, W, ch, cw, DEV, depth=2)
    pipe.submit(*pairs[0])
    for k in range(5):
        if k + 1 < 5:
            pipe.submit(*pairs[k + 1])
        left, right = pipe.next()
        wl, wr = IO.test_transform(IO.normalize_pair(*pairs[k]), ch, cw)
        assert float(np.abs(left.cpu().numpy() - wl).max()) <= 2e-6 and float(np.abs(right.cpu().numpy() - wr).max()) <= 2e-6
