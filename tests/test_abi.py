"""CPU tests (-m "not gpu"): the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/leastereo_b200.h declares; the product binding refuses to run the hot path without CUDA."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, seeded_model


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    g.build()
    return g.LIB


def test_header_symbols_exported(lib_path):
    header = open(os.path.join(ROOT, "include", "leastereo_b200.h")).read()
    declared = set(re.findall(r"\b(lea_[a-z0-9_]+)\s*\(", header))
    declared -= {"lea_vol", "lea_conv", "lea_tc_opts"}
    assert len(declared) >= 14
    lib = ctypes.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export " + name
    from leastereo_b200.kernels import SYMBOLS
    assert declared == set(SYMBOLS.keys()), declared ^ set(SYMBOLS.keys())


def test_library_identity(lib_path):
    lib = ctypes.CDLL(lib_path)
    assert lib.lea_abi_version() == 3
    assert lib.lea_is_device_build() == 1


def test_sass_is_blackwell_native(lib_path):
    import shutil, subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in sass, "TMA tensor load missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    # the forward / data-gradient path must be tcgen05 only; the one kernel allowed to use the warp-level tensor path
    # (mma.sync -> HMMA) is the weight gradient, whose tiny per-tap outputs do not fit tcgen05's 128 x N x 16 shape
    # (leastereo_b200/csrc/lea_wgrad_mma.cu)
    legacy = set()
    for chunk in sass.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0].strip()
        if "HMMA." in chunk.replace("UTCHMMA", ""):
            legacy.add(name)
    assert legacy and all("lea_wgrad_mma_kernel" in n for n in legacy), legacy
    conv = [c for c in sass.split("Function : ")[1:] if "lea_conv_tc_kernel" in c.split("\n", 1)[0]]
    assert conv and all("UTCHMMA" in c and "LDTM" in c for c in conv)


def test_argument_validation_without_gpu(lib_path):
    from leastereo_b200.kernels import Ops, lea_vol
    ops = Ops(lib_path, require_device_build=True)
    v = lea_vol(0, 1, 12, 2, 4, 4, 4)      # null data, channels not a multiple of 8
    rc = ops.lib.lea_trilinear_ac(ctypes.byref(v), 0, ctypes.byref(v), 0, 8, None, None, 0, None)
    assert rc != 0 and b"null volume" in ops.lib.lea_last_error()
    assert ops.tc_weight_image_bytes(64, 32, 3, 2) == 4 * 27 * 2 * 64 * 16
    assert ops.tc_weight_image_bytes(12, 32, 3, 2) == 0


def test_no_cpu_fallback():
    model = seeded_model(24).eval()
    x = torch.randn(1, 3, 24, 48)
    with pytest.raises(Exception) as ei:
        with torch.no_grad():
            model(x, x)
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_training_mode_fails_loudly():
    model = seeded_model(24).train()
    with pytest.raises(NotImplementedError):
        model.matching(torch.randn(1, 64, 8, 8, 16))
