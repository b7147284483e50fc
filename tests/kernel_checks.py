"""Parity checks of every kernel against the oracle, written once and run twice:
  * tests/test_gpu_kernels.py   (-m gpu)      through the real C-ABI library on a B200;
  * tests/test_emu_kernels.py   (-m "not gpu") through the CPU emulation build of the same kernel source, so the index
    math is verified in the GPU-less build container.
The checker is always the oracle (oracle/leastereo_oracle.py) or the golden fixtures; tolerances are stated inline.
"""
import hashlib

import numpy as np
import torch
import torch.nn.functional as F

from conftest import golden_arch, golden_genotype, golden_model, golden_state_dict, load_golden, seeded_model
from oracle import leastereo_oracle as O
from leastereo_b200.kernels import PlanesVol
from leastereo_b200 import engine


def _rand(shape, seed, device, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device)


# ---------------------------------------------------------------------------------------------------------
def check_cost_volume_f32(ops, device):
    # (B, C, H, W, maxdisp): vector path, scalar path (W % 4 != 0), D3 > W, D3 == 0
    for (B, Cn, H, W, maxdisp) in [(2, 8, 5, 16, 12), (1, 3, 4, 10, 9), (1, 8, 3, 4, 30), (1, 8, 3, 8, 2)]:
        x, y = _rand((B, Cn, H, W), 1, device), _rand((B, Cn, H, W), 2, device)
        got = ops.cost_volume_f32(x, y, maxdisp).cpu().numpy()
        want = O.cost_volume_numpy(x.cpu().numpy(), y.cpu().numpy(), maxdisp)
        assert got.shape == want.shape
        assert got.tobytes() == want.tobytes(), "cost volume must be bit-exact"


def check_cost_volume_golden(ops, device, name="cal_48x96_d48"):
    g = load_golden(name)
    fx, fy = torch.from_numpy(g["fx"]).to(device), torch.from_numpy(g["fy"]).to(device)
    cost = ops.cost_volume_f32(fx, fy, int(g["maxdisp"])).cpu().numpy()
    assert hashlib.sha256(cost.tobytes()).hexdigest() == str(g["cost_sha256"])


def check_pack_unpack(ops, device):
    x = _rand((2, 16, 3, 5, 7), 3, device, scale=100.0)
    x[0, 0, 0, 0, 0] = 0.0
    x[0, 1, 0, 0, 0] = 1e-30
    x[0, 2, 0, 0, 0] = -3e30
    v3 = ops.pack(x, 3)
    assert torch.equal(ops.unpack(v3), x), "3 planes must hold fp32 exactly"
    v2 = ops.pack(x, 2)
    err = (ops.unpack(v2) - x).abs() / x.abs().clamp_min(1e-30)
    assert float(err.max()) <= 2.0 ** -16, "2 planes keep 16 significant bits"
    v1 = ops.pack(x, 1)
    assert torch.equal(ops.unpack(v1), x.bfloat16().float())
    # channel-slice addressing
    big = PlanesVol.empty(2, 32, 3, 3, 5, 7, device)
    big.t.zero_()
    ops.pack(x, 3, out=big, c0=8)
    assert torch.equal(ops.unpack(big, 8, 16), x)
    assert float(ops.unpack(big, 0, 8).abs().max()) == 0.0 and float(ops.unpack(big, 24, 8).abs().max()) == 0.0


def check_cost_volume_planes(ops, device):
    x, y = _rand((2, 8, 6, 12), 4, device), _rand((2, 8, 6, 12), 5, device)
    vol = ops.cost_volume_planes(x, y, 15, 3)
    assert torch.equal(ops.unpack(vol), ops.cost_volume_f32(x, y, 15))


def check_trilinear(ops, device):
    cases = [((8, 6, 12), (4, 3, 6)), ((4, 3, 6), (8, 6, 12)), ((5, 9, 7), (3, 5, 4)), ((3, 5, 4), (5, 9, 7)),
             ((4, 4, 4), (4, 4, 4)), ((2, 2, 2), (1, 1, 1))]
    for i, (src_sp, dst_sp) in enumerate(cases):
        x = _rand((2, 16) + src_sp, 10 + i, device)
        src = ops.pack(x, 3)
        dst = PlanesVol.empty(2, 24, 3, *dst_sp, device)
        dst.t.zero_()
        ops.trilinear_ac(src, 8, 8, dst, 16)          # channels 8..15 -> slot 16..23
        got = ops.unpack(dst, 16, 8).cpu()
        want = F.interpolate(x[:, 8:16].cpu(), dst_sp, mode="trilinear", align_corners=True)
        assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max())), (src_sp, dst_sp)


def check_resample_conv1x1(ops, device):
    """lea_resample_conv1x1 against interpolate(align_corners=True) -> 1x1x1 conv -> BN -> ReLU for one and two consumers."""
    for i, (src_sp, dst_sp, n_out) in enumerate([((8, 6, 12), (4, 3, 6), 2), ((5, 9, 7), (3, 5, 4), 1), ((4, 4, 8), (2, 2, 4), 2)]):
        B, ct, c0, ci = 2, 24, 8, 16
        x = _rand((B, ct) + src_sp, 90 + i, device)
        src = ops.pack(x, 3)
        outs, refs = [], []
        xr = F.interpolate(x[:, c0:c0 + ci].cpu(), dst_sp, mode="trilinear", align_corners=True)
        for k, co in enumerate([8, 16][:n_out]):
            w = _rand((co, ci, 1, 1, 1), 95 + i + k, device, scale=0.3).contiguous()
            sc = (_rand((co,), 97 + k, device).abs() + 0.5).contiguous()
            sh = _rand((co,), 98 + k, device).contiguous()
            relu = (k == 0)
            dst = PlanesVol.empty(B, co + 8, 3, *dst_sp, device)
            dst.t.zero_()
            outs.append((dst, 8, co, w, sc, sh, relu))
            r = F.conv3d(xr, w.cpu()) * sc.cpu().view(1, -1, 1, 1, 1) + sh.cpu().view(1, -1, 1, 1, 1)
            refs.append(F.relu(r) if relu else r)
        ops.resample_conv1x1(src, c0, ci, outs)
        for (dst, dc0, co, *_), ref in zip(outs, refs):
            got = ops.unpack(dst, dc0, co).cpu()
            assert float((got - ref).abs().max()) <= 5e-6 * max(1.0, float(ref.abs().max())), (src_sp, dst_sp)
            assert float(ops.unpack(dst, 0, 8).abs().max()) == 0.0


def _conv_case(ops, device, B, c_in_total, c0, c_in, c_out, k, sp, bn, relu, res, planes, seed, conv_fn):
    x = _rand((B, c_in_total) + sp, seed, device)
    w = _rand((c_out, c_in, k, k, k), seed + 1, device, scale=0.2)
    scale = shift = None
    if bn:
        scale = (_rand((c_out,), seed + 2, device).abs() + 0.5).contiguous()
        shift = _rand((c_out,), seed + 3, device).contiguous()
    src = ops.pack(x, planes)
    ref = F.conv3d(x[:, c0:c0 + c_in].cpu(), w.cpu(), None, 1, (k - 1) // 2)
    if bn:
        ref = ref * scale.cpu().view(1, -1, 1, 1, 1) + shift.cpu().view(1, -1, 1, 1, 1)
    if relu:
        ref = F.relu(ref)
    if c_out % 8 == 0:
        dst = PlanesVol.empty(B, c_out + 16, planes, *sp, device)
        dst.t.zero_()
        r = None
        if res:
            r = _rand((B, c_out) + sp, seed + 4, device)
            ops.pack(r, planes, out=dst, c0=8)
            ref = ref + ops.unpack(dst, 8, c_out).cpu()
        p = ops.make_conv(src, c0, c_in, c_out, k, scale, shift, relu, dst=dst, dst_c0=8,
                          res=dst if res else None, res_c0=8)
        conv_fn(p, w, x)
        got = ops.unpack(dst, 8, c_out).cpu()
        assert float(ops.unpack(dst, 0, 8).abs().max()) == 0.0, "wrote outside its channel slice"
    else:
        out = torch.empty((B, c_out) + sp, dtype=torch.float32, device=device)
        p = ops.make_conv(src, c0, c_in, c_out, k, scale, shift, relu, dst_f32=out)
        conv_fn(p, w, x)
        got = out.cpu()
    return got, ref


def check_conv_simt(ops, device):
    fn = lambda p, w, x: ops.conv3d_simt(p, w.contiguous(), x)
    # (B, c_in_total, c0, c_in, c_out, k, spatial, bn, relu, res)
    cases = [
        (1, 16, 0, 16, 16, 3, (3, 16, 8), True, True, False),
        (2, 24, 8, 16, 8, 3, (2, 18, 10), True, True, True),      # partial tiles, channel slice, residual
        (1, 8, 0, 8, 32, 3, (4, 5, 3), False, False, False),
        (1, 32, 0, 32, 1, 3, (3, 6, 9), False, False, False),      # last_3-like: c_out = 1, fp32 output
        (1, 64, 0, 64, 64, 3, (2, 4, 8), True, True, False),
        (2, 32, 8, 16, 16, 1, (3, 17, 9), True, True, False),      # 1x1x1
        (1, 128, 0, 128, 32, 1, (2, 4, 8), True, True, True),
        (1, 8, 0, 8, 8, 1, (1, 1, 1), False, False, False),
    ]
    for i, (B, ct, c0, ci, co, k, sp, bn, relu, res) in enumerate(cases):
        got, ref = _conv_case(ops, device, B, ct, c0, ci, co, k, sp, bn, relu, res, 3, 100 + 10 * i, fn)
        tol = 2e-5 * max(1.0, float(ref.abs().max()))       # fp32 accumulate, different summation order
        assert float((got - ref).abs().max()) <= tol, (i, float((got - ref).abs().max()), tol)


def check_disp_head(ops, device):
    for (B, D3, H3, W3, maxdisp, seed) in [(2, 8, 5, 7, 24, 1), (1, 16, 4, 8, 50, 2), (1, 4, 9, 17, 12, 3)]:
        mat = _rand((B, 1, D3, H3, W3), seed, device, scale=3.0)
        got = ops.disp_head(mat, maxdisp).cpu()
        want = O.disp_head(mat.cpu(), maxdisp)
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= 2e-3, float((got - want).abs().max())   # px; __expf vs expf
    # raw-regime magnitudes (logits ~1e8): must not overflow, hard arg-min behaviour
    mat = _rand((1, 1, 8, 4, 8), 9, device, scale=1e8)
    got = ops.disp_head(mat, 24).cpu()
    want = O.disp_head(mat.cpu(), 24)
    assert torch.isfinite(got).all()
    assert float(((got - want).abs() <= 0.1).float().mean()) >= 0.99
    # un-normalised logits with a trend along disparity (what random-init weights produce, SURVEY App. C "raw" regime):
    # the minimum sits at an end of the range, where outputs 0/1 (and maxdisp-2/-1) read the same sample and the
    # reference's fp32 interpolation weights (l1 = 1.49e-8 at output 1) break the tie - the head must use those weights
    for sign, seed in ((1.0, 21), (-1.0, 22)):
        D3, H3, W3 = 16, 6, 40
        ramp = torch.arange(D3, dtype=torch.float32).view(1, 1, D3, 1, 1) * 2.0e7 * sign
        mat = (ramp + _rand((2, 1, D3, H3, W3), seed, "cpu", scale=3.0e6)).to(device)
        got = ops.disp_head(mat, 3 * D3).cpu()
        want = O.disp_head(mat.cpu(), 3 * D3)
        diff = (got - want).abs()
        assert float((diff <= 0.1).float().mean()) >= 0.999 and float(diff.mean()) <= 0.01, (sign, float(diff.max()))
        if sign > 0:        # the tie is really broken by the weight bits: neither 0.5 (exact thirds) nor 0 or 1
            assert 0.05 < float(want.median()) < 0.495, float(want.median())


def check_head_taps(ops, device):
    """lea_head_taps (+ the 1x1x1 tap projection) against  conv3d(interpolate(x, align_corners=True), w3, padding=1)
    - retrain/skip_model_3d.py:162-169 - on up-sampling shapes incl. odd sizes, a kept axis and a size-1 axis."""
    cases = [((4, 6, 8), (8, 12, 16)), ((3, 5, 4), (5, 9, 7)), ((4, 6, 8), (8, 6, 16)), ((1, 3, 5), (1, 6, 10)),
             ((2, 2, 2), (7, 9, 5))]
    for i, (src_sp, dst_sp) in enumerate(cases):
        B, Cn = 2, 16
        x = _rand((B, Cn) + src_sp, 40 + i, device)
        w3 = _rand((1, Cn, 3, 3, 3), 60 + i, device, scale=0.3)
        q = torch.einsum("bcdhw,ct->btdhw", x, w3.reshape(Cn, 27))
        qv = PlanesVol.empty(B, 40, 3, *src_sp, device)
        qv.t.zero_()
        ops.pack(F.pad(q, (0, 0, 0, 0, 0, 0, 0, 5)).contiguous(), 3, out=qv, c0=8)      # taps at channels 8..34
        mat = torch.empty((B, 1) + dst_sp, dtype=torch.float32, device=device)
        ops.head_taps(qv, 8, mat)
        want = F.conv3d(F.interpolate(x.cpu(), dst_sp, mode="trilinear", align_corners=True), w3.cpu(), None, 1, 1)
        err = float((mat.cpu() - want).abs().max()) / max(1.0, float(want.abs().max()))
        assert err <= 5e-6, (src_sp, dst_sp, err)


def check_stem0_collapse(ops, device):
    """Collapsed stem0: L/A/B 2-D maps (derived weights, ordinary conv kernel on depth-1 volumes) + lea_stem0_assemble
    against conv3d over the materialised cost volume (LEAStereo.py:34-48 + skip_model_3d.py:141) on every voxel the
    assemble kernel claims; the other voxels must stay untouched."""
    worst = 0.0
    # (B, maxdisp, H, W): D = 13 with a fully masked tile; ragged width (last tile partial); minimum depth 3;
    # D3 == W3 (every disparity column in play, deep masked triangle)
    for (B, maxdisp, H, W) in [(2, 39, 5, 40), (1, 30, 3, 27), (1, 9, 4, 24), (1, 96, 2, 32)]:
        worst = max(worst, _stem0_collapse_case(ops, device, B, maxdisp, H, W))
    return worst


def _stem0_collapse_case(ops, device, B, maxdisp, H, W):
    fm, co = 8, 8
    D = maxdisp // 3
    fx, fy = _rand((B, fm, H, W), 71 + W, device), _rand((B, fm, H, W), 72 + W, device)
    w = _rand((co, 2 * fm, 3, 3, 3), 73 + W, device, scale=0.2)
    scale = (_rand((co,), 74, device).abs() + 0.5).contiguous()
    shift = _rand((co,), 75, device).contiguous()
    wl, wab = engine.collapsed_stem0_weights(w, fm)
    fx3, fy3 = ops.pack(fx, 3), ops.pack(fy, 3)
    lmap = PlanesVol.empty(B, co, 3, 1, H, W, device)
    abmap = PlanesVol.empty(B, 2 * co, 3, 1, H, W, device)
    ops.conv3d_simt(ops.make_conv(fx3, 0, fm, co, 3, None, None, False, dst=lmap), wl.contiguous(), fx)
    ops.conv3d_simt(ops.make_conv(fy3, 0, fm, 2 * co, 3, None, None, False, dst=abmap), wab.contiguous(), fy)
    dst = PlanesVol.empty(B, co + 8, 3, D, H, W, device)
    dst.t.zero_()
    ops.stem0_assemble(lmap, abmap, dst, 8, co, scale, shift, True)
    got = ops.unpack(dst, 8, co).cpu()
    cost = torch.from_numpy(O.cost_volume_numpy(fx.cpu().numpy(), fy.cpu().numpy(), maxdisp))
    ref = F.relu(F.conv3d(cost, w.cpu(), None, 1, 1) * scale.cpu().view(1, -1, 1, 1, 1) + shift.cpu().view(1, -1, 1, 1, 1))
    claimed = torch.zeros((D, W), dtype=torch.bool)
    for d in range(D):
        for x in range(W):
            claimed[d, x] = ((1 <= d <= D - 2) and (8 * (x // 8) >= d + 2) and (8 * (x // 8) + 7 <= W - 2)) or \
                (8 * (x // 8) + 7 <= d - 3)
    assert int(claimed.sum()) > 0, (D, W)
    m = claimed.view(1, 1, D, 1, W).expand_as(got)
    err = float((got - ref)[m].abs().max()) / float(ref.abs().max())
    assert err <= 2e-6, err
    assert float(got[~m].abs().max()) == 0.0, "wrote outside the claimed voxels"
    assert float(ops.unpack(dst, 0, 8).abs().max()) == 0.0
    return err


def check_stem0_collapse_tc(ops, device):
    """GPU only: the two kernels that share stem0's output - lea_stem0_assemble and the tensor-core launch with the
    fused cost-volume loader and cv_skip - must together write EVERY voxel exactly once and agree with conv3d over the
    materialised cost volume; ragged width, a depth range with masked tiles, batch 2."""
    from leastereo_b200.kernels import lea_tc_opts
    worst = 0.0
    for (B, maxdisp, H, W) in [(2, 39, 20, 43), (1, 75, 16, 32), (1, 9, 16, 24)]:
        fm, co, P = 16, 16, 2
        D = maxdisp // 3
        fx, fy = _rand((B, fm, H, W), 81 + W, device), _rand((B, fm, H, W), 82 + W, device)
        w = _rand((co, 2 * fm, 3, 3, 3), 83 + W, device, scale=0.2)
        scale = (_rand((co,), 84, device).abs() + 0.5).contiguous()
        shift = _rand((co,), 85, device).contiguous()
        # 2-D maps (exact planes, fp32 FMA convs) + assemble
        wl, wab = engine.collapsed_stem0_weights(w, fm)
        fx3, fy3 = ops.pack(fx, 3), ops.pack(fy, 3)
        lmap = PlanesVol.empty(B, co, 3, 1, H, W, device)
        abmap = PlanesVol.empty(B, 2 * co, 3, 1, H, W, device)
        ops.conv3d_simt(ops.make_conv(fx3, 0, fm, co, 3, None, None, False, dst=lmap), wl.contiguous(), fx)
        ops.conv3d_simt(ops.make_conv(fy3, 0, fm, 2 * co, 3, None, None, False, dst=abmap), wab.contiguous(), fy)
        dst = PlanesVol.empty(B, co, P, D, H, W, device)
        dst.t.fill_(float("nan"))                              # every voxel must be overwritten
        ops.stem0_assemble(lmap, abmap, dst, 0, co, scale, shift, True)
        # band + edges: tensor-core kernel, cost volume built by its loader, skipping the assembled voxels
        fxy = PlanesVol.empty(2 * B, fm, P, 1, H, W, device)
        fxp, fyp = PlanesVol(fxy.t[:B]), PlanesVol(fxy.t[B:])
        ops.pack(fx, P, out=fxp); ops.pack(fy, P, out=fyp)
        maps = ops.build_fused_cv_maps(fxp, fyp, D)
        p = ops.make_conv(fxp, 0, 2 * fm, co, 3, scale, shift, True, dst=dst, dst_c0=0)
        opts = lea_tc_opts()
        opts.fused_cv, opts.fx, opts.fy, opts.d3 = 1, fxp.struct(), fyp.struct(), D
        opts.cv_maps = maps.data_ptr()
        opts.cv_skip = 1
        ops.conv3d_tc(p, ops.pack_weights_tc(w.contiguous(), P), opts, fx)
        torch.cuda.synchronize()
        got = ops.unpack(dst).cpu()
        assert torch.isfinite(got).all(), "a voxel was written by neither kernel"
        cost = torch.from_numpy(O.cost_volume_numpy(fx.cpu().numpy(), fy.cpu().numpy(), maxdisp))
        ref = F.relu(F.conv3d(cost, w.cpu(), None, 1, 1) * scale.cpu().view(1, -1, 1, 1, 1) + shift.cpu().view(1, -1, 1, 1, 1))
        err = float((got - ref).abs().max()) / float(ref.abs().max())
        assert err <= 2e-4, (B, maxdisp, H, W, err)            # 2-plane operands / storage
        worst = max(worst, err)
    return worst


def check_disparity_regression(ops, device):
    p = torch.softmax(_rand((2, 24, 5, 7), 5, device), dim=1).contiguous()
    got = ops.disparity_regression(p, 24).cpu()
    want = (p.cpu() * torch.arange(24.0).view(1, 24, 1, 1)).sum(1)
    assert float((got - want).abs().max()) <= 1e-4


def run_hot_path(ops, device, g, planes, conv="simt", mma_terms=0, extra=None):
    """Product engine on a golden case's feature maps; returns (mat, disp) on CPU."""
    maxdisp = int(g["maxdisp"])
    model = golden_model(g, device).eval()
    model.engine_options = {"planes": planes, "conv": conv, "mma_terms": mma_terms}
    model.engine_options.update(extra or {})
    fx, fy = torch.from_numpy(g["fx"]).to(device), torch.from_numpy(g["fy"]).to(device)
    disp = engine.hot_path_forward(model, fx, fy, ops=ops)
    B, _, H3, W3 = fx.shape
    plan = engine.get_plan(model.matching, B, (int(maxdisp / 3), H3, W3), fx.device, engine._options(model), ops)
    return plan.mat.cpu().clone(), disp.cpu(), model, plan


def check_hot_path_golden(ops, device, name, planes, conv="simt", mma_terms=0, mat_rtol=2e-4, require_tolerance=True,
                          extra=None):
    g = load_golden(name)
    mat, disp, _, _ = run_hot_path(ops, device, g, planes, conv, mma_terms, extra)
    ref_mat, ref_disp = torch.from_numpy(g["mat"]), torch.from_numpy(g["disp"])
    assert mat.shape == ref_mat.shape and disp.shape == ref_disp.shape
    rel = float((mat - ref_mat).abs().max()) / float(ref_mat.abs().max())
    rep = O.tolerance_report(disp, ref_disp)
    rep["mat_rel_err"] = rel
    if mat_rtol is not None:
        assert rel <= mat_rtol, rep
    if require_tolerance:
        # north-star tolerance: |dd| <= 0.1 px on >= 99.9 % of pixels, mean |dd| <= 0.01 px
        assert rep["ok"], rep
    return rep


# ---------------------------------------------------------------------------------------------------------
# tcgen05 conv (GPU only)
# ---------------------------------------------------------------------------------------------------------
TC_CASES = [
    # (B, c_in_total, c0, c_in, c_out, k, spatial, bn, relu, res)
    (1, 16, 0, 16, 16, 3, (4, 16, 8), False, False, False),      # one full tile, smallest K
    (1, 16, 0, 16, 16, 3, (3, 16, 8), True, True, False),
    (2, 48, 16, 32, 32, 3, (5, 18, 10), True, True, True),       # partial tiles, slice, residual, 2 channel groups
    (1, 64, 0, 64, 32, 3, (6, 32, 24), True, True, False),       # stem0-like
    (1, 128, 0, 128, 64, 3, (4, 16, 16), True, True, False),     # conv1-like (single weight buffer)
    (1, 8, 0, 8, 8, 3, (4, 16, 16), True, True, True),           # 8-channel layout (cell 10)
    (1, 32, 0, 32, 1, 3, (5, 16, 8), False, False, False),       # last_3: c_out = 1, fp32 output
    (2, 32, 0, 32, 16, 1, (3, 17, 9), True, True, False),        # 1x1x1
    (1, 128, 0, 128, 32, 1, (9, 16, 8), True, True, True),
    (1, 64, 0, 64, 8, 1, (2, 16, 16), True, True, False),
    (1, 32, 0, 32, 32, 3, (19, 40, 24), True, True, False),      # several work items per CTA, depth chunks
]


# forced depth-chunk lengths (last field): several work items per tile column, short last chunks, chunks longer than
# the volume, one and two accumulator sets
TC_CHUNK_CASES = [
    (1, 32, 0, 32, 32, 3, (19, 20, 12), True, True, False, 19),
    (1, 32, 0, 32, 32, 3, (19, 20, 12), True, True, True, 7),       # chunks 7+7+5
    (2, 16, 0, 16, 16, 3, (40, 16, 8), True, True, True, 40),
    (1, 16, 0, 16, 16, 3, (33, 16, 8), False, False, False, 15),
    (1, 8, 0, 8, 8, 3, (70, 16, 8), True, True, True, 70),
    (1, 64, 0, 64, 32, 3, (13, 16, 8), True, True, False, 13),      # weights streamed per slab and channel group
    (1, 32, 0, 32, 1, 3, (21, 16, 8), False, False, False, 21),     # fp32 output (per-thread stores)
    (1, 16, 0, 16, 48, 3, (6, 16, 8), True, True, False, 6),        # 96 columns per depth, two sets
    (1, 128, 0, 128, 64, 3, (9, 16, 16), True, True, False, 4),     # streamed weights, 8 channel groups
    (1, 16, 0, 16, 24, 3, (9, 21, 13), True, True, True, 4),        # 16-channel group + 8-channel tail, ragged tiles
    (2, 64, 16, 32, 16, 1, (5, 9, 70), True, True, True, 3),        # 1x1x1, wide tile (32 x 4), clipped rows and columns
    (1, 32, 0, 32, 8, 1, (3, 20, 24), True, False, False, 2),       # 1x1x1, 16 x 8 tile
]


def check_conv_tc(ops, device, planes=2, mma_terms=0, cases=None, verbose=False):
    from leastereo_b200.kernels import lea_tc_opts
    worst = 0.0
    for i, case in enumerate(cases or TC_CASES):
        (B, ct, c0, ci, co, k, sp, bn, relu, res), chunk = case[:10], (case[10] if len(case) > 10 else 0)
        if ops.tc_weight_image_bytes(ci, co, k, planes) <= 0:
            continue

        def fn(p, w, x):
            img = ops.pack_weights_tc(w.contiguous(), planes)
            opts = lea_tc_opts()
            opts.mma_terms = mma_terms
            opts.depth_chunk = abs(chunk)
            ops.conv3d_tc(p, img, opts, x)

        got, ref = _conv_case(ops, device, B, ct, c0, ci, co, k, sp, bn, relu, res, planes, 300 + 10 * i, fn)
        torch.cuda.synchronize()
        scale = max(1.0, float(ref.abs().max()))
        err = float((got - ref).abs().max()) / scale
        # error model: operands carry 8 bits per plane; dropped cross terms ~2^-(8*planes) of |a||w| per product
        tol = {1: 3e-2, 2: 2e-4, 3: 2e-5}[planes] if mma_terms == 0 else 3e-2
        if verbose:
            print("conv_tc case %d planes %d terms %d: rel err %.3g (tol %.1g)" % (i, planes, mma_terms, err, tol))
        assert err <= tol, (i, planes, mma_terms, err, tol)
        worst = max(worst, err)
    return worst


def check_wgrad_tc(ops, device, verbose=False):
    """lea_conv3d_wgrad_tc (mma.sync on the two bf16 planes, bf16x3 products) against the weight gradient of F.conv3d
    computed in fp64 on the SAME 2-plane-rounded operands; ragged tiles, channel slices, every channel tiling
    (8 / 16 / 32), several depth chunks, and against the fp32 FMA kernel it replaces.
    Tolerance 3e-5 of the gradient's max: the dropped lo*lo products (2^-16 per product, random sign)."""
    worst = 0.0
    cases = [  # (B, c_in_total, c0, c_in, c_out_total, o0, c_out, spatial[, ksize])
        (2, 64, 0, 64, 16, 0, 16, (5, 19, 13), 1),       # 1x1x1: cells' pre-processing layers
        (1, 128, 0, 128, 32, 0, 32, (3, 16, 24), 1),
        (1, 64, 0, 64, 8, 0, 8, (4, 9, 30), 1),
        (2, 48, 16, 32, 24, 8, 16, (6, 17, 8), 1),
        (1, 8, 0, 8, 8, 0, 8, (5, 16, 8)),
        (2, 16, 0, 16, 16, 0, 16, (7, 19, 13)),
        (1, 32, 0, 32, 32, 0, 32, (9, 33, 17)),
        (1, 64, 0, 64, 32, 0, 32, (4, 16, 24)),          # two ci tiles (stem0-like)
        (1, 96, 32, 64, 72, 8, 64, (3, 20, 9)),          # channel slices, 2 x 2 tiles of 32
        (2, 16, 8, 8, 24, 16, 8, (6, 9, 30)),            # 8-channel tiles inside wider volumes
        (1, 48, 0, 48, 16, 0, 16, (3, 16, 16)),          # 3 ci tiles of 16
        (1, 16, 0, 16, 16, 0, 16, (70, 16, 8)),          # long column: several depth chunks / ring wraps
    ]
    for i, case in enumerate(cases):
        (B, cit, c0, ci, cot, o0, co, sp), k = case[:8], (case[8] if len(case) > 8 else 3)
        assert ops.lib.lea_conv3d_wgrad_tc_supported(ci, co, k, 2)
        x = _rand((B, cit) + sp, 500 + i, device)
        dy = _rand((B, cot) + sp, 600 + i, device)
        xv, dyv = ops.pack(x, 2), ops.pack(dy, 2)
        dw = torch.zeros((co, ci, k, k, k), dtype=torch.float32, device=device)
        ops.conv3d_wgrad(xv, c0, ci, dyv, o0, co, k, dw, tensor_cores=True)
        dw_fma = torch.zeros_like(dw)
        ops.conv3d_wgrad(xv, c0, ci, dyv, o0, co, k, dw_fma, tensor_cores=False)
        xr = ops.unpack(xv).cpu().double()[:, c0:c0 + ci]                 # the operands the kernels actually see
        dr = ops.unpack(dyv).cpu().double()[:, o0:o0 + co]
        w = torch.zeros((co, ci, k, k, k), dtype=torch.float64, requires_grad=True)
        (F.conv3d(xr, w, None, 1, (k - 1) // 2) * dr).sum().backward()
        scale = float(w.grad.abs().max())
        err = float((dw.cpu().double() - w.grad).abs().max()) / scale
        err_fma = float((dw_fma.cpu().double() - w.grad).abs().max()) / scale
        if verbose:
            print("wgrad_tc case %d: rel err %.3g (fma kernel %.3g)" % (i, err, err_fma))
        assert err <= 3e-5, (i, err)
        worst = max(worst, err)
    return worst


# ---------------------------------------------------------------------------------------------------------
# native feature net (2D net on depth-1 planes volumes + fused stems) against the stock-PyTorch module
# ---------------------------------------------------------------------------------------------------------
def check_feature_plan(ops, device, name="cal_46x94_d50", planes=3, conv="simt", tol=2e-4):
    g = load_golden(name)
    model = seeded_model(int(g["maxdisp"]))
    model.load_state_dict(golden_state_dict(g, model))
    model = model.to(device).eval()
    left = torch.from_numpy(g["left"]).to(device)
    right = torch.from_numpy(g["right"]).to(device)
    B, _, H, W = left.shape
    h3, w3 = (H - 1) // 3 + 1, (W - 1) // 3 + 1
    out = PlanesVol.empty(2 * B, 32, planes, 1, h3, w3, device)
    plan = engine.FeaturePlan(model.feature, ops, 2 * B, H, W, planes, device, out, conv_mode=conv)
    plan.img[:B].copy_(left)
    plan.img[B:].copy_(right)
    plan.run()
    got = ops.unpack(out)[:, :, 0].cpu()
    want = torch.cat([torch.from_numpy(g["fx"]), torch.from_numpy(g["fy"])], dim=0)     # the reference's feature maps
    err = float((got - want).abs().max()) / float(want.abs().max())
    assert got.shape == want.shape
    assert err <= tol, err
    return err


# ---------------------------------------------------------------------------------------------------------
# callers on either side of the path: input normalisation, loss, Adam, metrics
# ---------------------------------------------------------------------------------------------------------
def check_normalize_pad(ops, device):
    """lea_image_stats_u8 + lea_normalize_pad_u8 against (i) the outputs of predict.py's own load_data /
    test_transform (tests/golden/io_predict.npz, written by tests/golden/make_io_golden.py) and (ii) the restatement in
    oracle/io_oracle.py.  Tolerance 2e-6: the reference rounds (x - mean) / std once to float32 from numpy's fp64
    mean / std, the kernel from exact integer sums - the last-ulp difference of a z-score of magnitude <= 4."""
    from oracle import io_oracle as IO
    z = load_golden("io_predict")
    for i in range(int(z["n_cases"])):
        ch, cw = (int(v) for v in z["crop%d" % i])
        for img, want in ((z["left%d" % i], z["input1_%d" % i]), (z["right%d" % i], z["input2_%d" % i])):
            got = ops.normalize_pad_u8(torch.from_numpy(img).to(device), ch, cw).cpu().numpy()
            assert got.shape == want.shape[1:], i
            assert float(np.abs(got - want[0]).max()) <= 2e-6, i
    g = np.random.RandomState(7)
    for (h, w, ch, cw) in [(20, 30, 24, 36), (20, 30, 20, 30), (20, 30, 16, 24), (37, 53, 24, 48), (5, 300, 8, 512)]:
        left = g.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        right = g.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        left[:, :, 1] //= 4                                  # a low-contrast channel
        want_l, want_r = IO.test_transform(IO.normalize_pair(left, right), ch, cw)
        got_l = ops.normalize_pad_u8(torch.from_numpy(left).to(device), ch, cw).cpu().numpy()
        got_r = ops.normalize_pad_u8(torch.from_numpy(right).to(device), ch, cw).cpu().numpy()
        assert got_l.shape == want_l.shape[1:]
        assert float(np.abs(got_l - want_l[0]).max()) <= 2e-6, (h, w, ch, cw)
        assert float(np.abs(got_r - want_r[0]).max()) <= 2e-6, (h, w, ch, cw)


def _loss_case(device, seed=3, shape=(2, 17, 23), maxdisp=48.0):
    g = torch.Generator().manual_seed(seed)
    target = torch.rand(shape, generator=g) * maxdisp * 1.2            # some beyond maxdisp
    target[torch.rand(shape, generator=g) < 0.2] = 0.0                 # occlusions
    disp = target + torch.randn(shape, generator=g) * 1.5              # both smooth-L1 branches
    return disp.to(device), target.to(device), maxdisp


def check_masked_smooth_l1(ops, device):
    from oracle import io_oracle as IO
    disp, target, maxdisp = _loss_case(device)
    acc = ops.masked_smooth_l1(disp, target, maxdisp).cpu()
    d_ref = disp.detach().cpu().clone().requires_grad_(True)
    want = IO.masked_smooth_l1(d_ref, target.cpu(), maxdisp)
    want.backward()
    mask = (target.cpu() < maxdisp) & (target.cpu() > 0.001)
    assert int(acc[2]) == int(mask.sum())
    wv = float(want.detach())
    assert abs(float(acc[0] / acc[2]) - wv) <= 1e-6 * max(1.0, abs(wv))
    assert abs(float(acc[1] / acc[2]) - float((disp.cpu() - target.cpu())[mask].abs().mean())) <= 1e-5
    grad = ops.masked_smooth_l1_bwd(disp, target, maxdisp, acc.to(device), 1.0).cpu()
    assert float((grad - d_ref.grad).abs().max()) <= 1e-8
    # no valid pixel: zero gradient, no NaN
    z = torch.zeros_like(target)
    acc0 = ops.masked_smooth_l1(disp, z, maxdisp)
    assert float(acc0[2]) == 0.0
    assert float(ops.masked_smooth_l1_bwd(disp, z, maxdisp, acc0, 1.0).abs().max()) == 0.0


def check_adam(ops, device):
    """FlatAdam (one lea_adam_step launch over the flat buffer) against torch.optim.Adam (train.py:76)."""
    from leastereo_b200.pipeline import FlatAdam, multistep_lr
    g = torch.Generator().manual_seed(11)
    shapes = [(4, 3, 3, 3, 3), (7,), (5, 6)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g).to(device)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ours]
    opt = FlatAdam(ours, lr=1e-3, betas=(0.9, 0.999), ops=ops)
    ropt = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.999))
    for step in range(6):
        opt.zero_grad(); ropt.zero_grad()
        for p, r in zip(ours, ref):
            gr = torch.randn(p.shape, generator=g) * (10.0 ** (step - 3))
            r.grad = gr.clone()
            if step % 2 == 0:
                p.grad.copy_(gr.to(device))                  # accumulate into the flat bucket (what autograd does)
            else:
                p.grad = gr.to(device)                       # a foreign .grad tensor: gathered by step()
        opt.step(); ropt.step()
        for p, r in zip(ours, ref):
            assert float((p.detach().cpu() - r.detach()).abs().max()) <= 2e-6 * max(1.0, float(r.abs().max())), step
    assert multistep_lr(1e-3, 0, [30, 50]) == 1e-3 and multistep_lr(1e-3, 30, [30, 50]) == 5e-4
    assert multistep_lr(1e-3, 70, [30, 50]) == 2.5e-4


def check_disparity_metrics(ops, device):
    """lea_disparity_metrics against (i) the fixtures written by RUNNING utils/metrics.py / evaluation.py:290-292
    (tests/golden/io_metrics.npz) and (ii) the oracle restatement on a seeded case; counts are integers, so the
    fractions must agree to rounding (1e-12), the EPE (float32 mean in numpy, fp64 sum here) to 1e-5."""
    from oracle import io_oracle as IO
    from leastereo_b200.pipeline import disparity_metrics
    z = load_golden("io_metrics")
    for i in range(int(z["n_cases"])):
        p, t, maxdisp = z["pred%d" % i], z["true%d" % i], int(z["maxdisp%d" % i])
        want = z["result%d" % i]                                        # [3px, bad1, bad2, bad3, bad5, epe, #valid]
        got = disparity_metrics(torch.from_numpy(p).to(device), torch.from_numpy(t).to(device), maxdisp,
                                (1.0, 2.0, 3.0, 5.0), ops=ops)
        assert got["valid"] == int(want[6]), i
        assert abs(got["three_px_error"] - want[0]) <= 1e-12, (i, got, want)
        for k, thr in enumerate((1.0, 2.0, 3.0, 5.0)):
            assert abs(got["bad_%g" % thr] - want[1 + k]) <= 1e-12, (i, thr, got, want)
        assert abs(got["epe"] - want[5]) <= 1e-5 * max(1.0, abs(want[5])), (i, got, want)
    disp, target, maxdisp = _loss_case(device, seed=5, shape=(3, 31, 29), maxdisp=96.0)
    disp = disp + (torch.rand(disp.shape) < 0.1).to(device) * 7.0          # some gross errors
    got = disparity_metrics(disp, target, maxdisp, (1.0, 2.0, 3.0, 5.0), ops=ops)
    p, t = disp.cpu().numpy(), target.cpu().numpy()
    assert got["valid"] == int(IO.validity_mask(t, maxdisp).sum())
    assert abs(got["epe"] - IO.epe(p, t, maxdisp)) <= 1e-5
    assert abs(got["train_epe"] - IO.train_epe(p, t, maxdisp)) <= 1e-5
    assert abs(got["three_px_error"] - IO.three_px_error(p, t, maxdisp)) <= 1e-12
    for thr in (1.0, 2.0, 3.0, 5.0):
        assert abs(got["bad_%g" % thr] - IO.bad_pixel_frac(p, t, maxdisp, thr)) <= 1e-12
    # the opt-in un-truncated variant differs from the reference exactly as ADVICE r1 measured (more pixels "bad")
    flt = disparity_metrics(disp, target, maxdisp, (1.0, 2.0, 3.0, 5.0), ops=ops, float_diff=True)
    assert flt["bad_1"] > got["bad_1"] and flt["valid"] == got["valid"]
    mask = IO.validity_mask(t, maxdisp)
    assert abs(flt["bad_1"] - (1.0 - float((np.abs(t - p)[mask] <= 1.0).sum()) / float(mask.sum()))) <= 1e-12
