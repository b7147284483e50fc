"""CPU tests (-m "not gpu"): the SIMT kernel SOURCE of the product, compiled by g++ against tests/emu/cuda_emu.h
(thread-per-CUDA-thread emulation), checked against the oracle.  This verifies index math, layouts, the engine's
launch list and buffer wiring in the GPU-less build container; it is not a product path (leastereo_b200 never loads
the emulation library, see leastereo_b200/kernels.py:get_ops)."""
import os
import shutil
import subprocess

import pytest
import torch

import kernel_checks as K
from conftest import ROOT
from leastereo_b200.kernels import Ops

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libleastereo_emu.so")


@pytest.fixture(scope="session")
def emu_ops():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    srcs = [os.path.join(EMU_DIR, "emu_lib.cpp"), os.path.join(EMU_DIR, "cuda_emu.h")] + \
        [os.path.join(ROOT, "leastereo_b200", "csrc", f) for f in
         ("lea_common.h", "lea_simt_kernels.cuh", "lea_api_simt.inl", "lea_train_kernels.cuh", "lea_api_train.inl",
          "lea_io_kernels.cuh", "lea_api_io.inl")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in srcs):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-DLEA_CPU_EMU", "-Wno-unknown-pragmas",
                               "-I" + EMU_DIR, "-o", EMU_LIB, srcs[0], "-lpthread"])
    torch.set_num_threads(4)
    return Ops(EMU_LIB, require_device_build=False)


DEV = torch.device("cpu")


def test_emu_is_not_a_device_build(emu_ops):
    assert not emu_ops.device_build


def test_cost_volume_f32(emu_ops):
    K.check_cost_volume_f32(emu_ops, DEV)


def test_cost_volume_golden(emu_ops):
    K.check_cost_volume_golden(emu_ops, DEV, "cal_b2_24x48_d24")


def test_pack_unpack(emu_ops):
    K.check_pack_unpack(emu_ops, DEV)


def test_cost_volume_planes(emu_ops):
    K.check_cost_volume_planes(emu_ops, DEV)


def test_trilinear(emu_ops):
    K.check_trilinear(emu_ops, DEV)


def test_resample_conv1x1(emu_ops):
    K.check_resample_conv1x1(emu_ops, DEV)


def test_conv_simt(emu_ops):
    K.check_conv_simt(emu_ops, DEV)


def test_disp_head(emu_ops):
    K.check_disp_head(emu_ops, DEV)


def test_head_taps(emu_ops):
    K.check_head_taps(emu_ops, DEV)


def test_stem0_collapse(emu_ops):
    print("collapsed stem0 vs conv3d(cost volume): rel err", K.check_stem0_collapse(emu_ops, DEV))


def test_disparity_regression(emu_ops):
    K.check_disparity_regression(emu_ops, DEV)


def test_hot_path_golden_exact_planes(emu_ops):
    # 3 planes = fp32 storage, fp32 FMA convs: must sit at the fp32 re-ordering noise floor
    rep = K.check_hot_path_golden(emu_ops, DEV, "cal_b2_24x48_d24", planes=3, mat_rtol=2e-4)
    assert rep["max_abs"] <= 0.05, rep


def test_hot_path_skip_connect_genotype(emu_ops):
    """A 3D genotype with skip_connect ops (Identity inside the step sums) against the reference's own run of it; the
    plan must execute them as copy/accumulate passes, not as convolutions."""
    from conftest import load_golden
    g = load_golden("cal_skip_b2_24x48_d24")
    mat, disp, model, plan = K.run_hot_path(emu_ops, DEV, g, planes=3)
    copies = [s for s in plan.steps if s.kind == "copy"]
    assert len(copies) == 2 * 12, [s.name for s in copies]                 # two Identity ops in each of the 12 cells
    rep = K.check_hot_path_golden(emu_ops, DEV, "cal_skip_b2_24x48_d24", planes=3, mat_rtol=2e-4)
    print(rep)


def test_hot_path_golden_two_planes(emu_ops):
    # 2 planes (the bf16x3 operand format): activations carry 16 significant bits between layers
    rep = K.check_hot_path_golden(emu_ops, DEV, "cal_b2_24x48_d24", planes=2, mat_rtol=5e-3)
    print(rep)


def test_native_feature_net(emu_ops):
    # ragged image size (46x94): partial tiles, stride-3 stem with H, W not multiples of 3
    err = K.check_feature_plan(emu_ops, DEV, "cal_46x94_d50", planes=3, conv="simt")
    print("native feature net vs reference feature maps: rel err", err)


# ---- callers on either side of the path ------------------------------------------------
def test_normalize_pad(emu_ops):
    K.check_normalize_pad(emu_ops, DEV)


def test_masked_smooth_l1(emu_ops):
    K.check_masked_smooth_l1(emu_ops, DEV)


def test_flat_adam(emu_ops):
    K.check_adam(emu_ops, DEV)


def test_disparity_metrics(emu_ops):
    K.check_disparity_metrics(emu_ops, DEV)


def test_buffer_reuse_is_exact_and_smaller(emu_ops):
    """The liveness-packed arena must not change a single bit of the result and must shrink the activation memory."""
    from conftest import load_golden
    g = load_golden("cal_b2_24x48_d24")
    outs, sizes = {}, {}
    for reuse in (False, True):
        mat, disp, model, plan = K.run_hot_path(emu_ops, DEV, g, planes=2, extra={"reuse_buffers": reuse})
        outs[reuse] = (mat, disp)
        sizes[reuse] = plan.workspace_bytes()
        assert (plan._arena is not None) == reuse
    assert torch.equal(outs[True][0], outs[False][0]) and torch.equal(outs[True][1], outs[False][1])
    print("activation bytes: plain %d, reused %d" % (sizes[False], sizes[True]))
    assert sizes[True] < 0.6 * sizes[False]


def test_fused_resample_conv_is_on_the_plan_and_matches(emu_ops):
    """The down-sampling resample + 1x1x1 ConvBR fusion (lea_resample_conv1x1) replaces resample + conv launches and
    agrees with the two-launch form to the rounding of the intermediate it no longer stores."""
    from conftest import load_golden
    g = load_golden("cal_b2_24x48_d24")
    outs = {}
    for fused in (False, True):
        mat, disp, model, plan = K.run_hot_path(emu_ops, DEV, g, planes=3, extra={"fuse_resample_conv": fused})
        kinds = [s.kind for s in plan.steps]
        assert ("resample_conv1x1" in kinds) == fused
        outs[fused] = (mat, disp, len(kinds))
    assert outs[True][2] < outs[False][2]
    assert float((outs[True][1] - outs[False][1]).abs().max()) <= 1e-3
    assert float((outs[True][0] - outs[False][0]).abs().max()) <= 2e-4 * float(outs[False][0].abs().max())
