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
