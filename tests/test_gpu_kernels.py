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


def test_conv_simt(ops):
    K.check_conv_simt(ops, DEV)


def test_disp_head(ops):
    K.check_disp_head(ops, DEV)


def test_disparity_regression(ops):
    K.check_disparity_regression(ops, DEV)


@pytest.mark.parametrize("name", ["raw_48x96_d48", "cal_48x96_d48", "cal_46x94_d50", "cal_b2_24x48_d24"])
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


def test_conv_tc_single_pass(ops):
    K.check_conv_tc(ops, DEV, planes=2, mma_terms=1, verbose=True)


@pytest.mark.parametrize("name", ["raw_48x96_d48", "cal_48x96_d48", "cal_46x94_d50", "cal_b2_24x48_d24"])
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
        model.engine_options = {"planes": 2, "conv": "tc", "fuse_cv": fuse_cv}
        fx, fy = torch.from_numpy(g["fx"]).to(DEV), torch.from_numpy(g["fy"]).to(DEV)
        from leastereo_b200 import engine
        disp = engine.hot_path_forward(model, fx, fy, ops=ops)
        plan = engine.get_plan(model.matching, 1, (16, 16, 32), fx.device, engine._options(model), ops)
        assert (plan.fxp is not None) == fuse_cv
        outs[fuse_cv] = (plan.mat.clone(), disp.clone())
    assert torch.equal(outs[False][0], outs[True][0])
    assert torch.equal(outs[False][1], outs[True][1])
