"""CPU tests (-m "not gpu"): the oracle restatement against the golden vectors produced by the reference, and,
where /root/reference is mounted (build container only), against the live reference."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, REFERENCE_DIR, golden_arch, golden_state_dict, load_golden
from oracle import leastereo_oracle as O


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cost_volume_bit_exact(name):
    g = load_golden(name)
    cost = O.cost_volume_numpy(g["fx"], g["fy"], int(g["maxdisp"]))
    assert list(cost.shape) == list(g["cost_shape"])
    assert hashlib.sha256(cost.tobytes()).hexdigest() == str(g["cost_sha256"])
    assert np.array_equal(cost[:, ::7, ::3, ::5, ::3], g["cost_sample"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_hot_path_against_golden(name):
    g = load_golden(name)
    sd = golden_state_dict(g)
    maxdisp = int(g["maxdisp"])
    with torch.no_grad():
        cost = O.cost_volume(torch.from_numpy(g["fx"]), torch.from_numpy(g["fy"]), maxdisp)
        mat = O.matching_forward(sd, cost, golden_arch(g))
        disp = O.disp_head(mat, maxdisp)
    ref_mat = torch.from_numpy(g["mat"])
    assert mat.shape == ref_mat.shape
    scale = float(ref_mat.abs().max())
    assert float((mat - ref_mat).abs().max()) <= 1e-5 * scale
    rep = O.tolerance_report(disp, torch.from_numpy(g["disp"]))
    assert rep["max_abs"] <= 1e-3, rep


@pytest.mark.parametrize("name", ["cal_48x96_d48", "cal_46x94_d50"])
def test_feature_and_full_forward_against_golden(name):
    g = load_golden(name)
    sd = golden_state_dict(g)
    stages = {}
    disp = O.leastereo_forward(sd, torch.from_numpy(g["left"]), torch.from_numpy(g["right"]), int(g["maxdisp"]),
                               stages=stages)
    assert float((stages["fx"] - torch.from_numpy(g["fx"])).abs().max()) <= 1e-5
    assert float((stages["fy"] - torch.from_numpy(g["fy"])).abs().max()) <= 1e-5
    rep = O.tolerance_report(disp, torch.from_numpy(g["disp"]))
    assert rep["max_abs"] <= 1e-3, rep


def test_calibration_reproduces_golden_bn_stats():
    g = load_golden("cal_48x96_d48")
    raw = load_golden("raw_48x96_d48")
    sd0 = golden_state_dict(raw)   # plain seeded init
    sd = O.calibrate_bn(sd0, torch.from_numpy(g["left"]), torch.from_numpy(g["right"]), int(g["maxdisp"]))
    n = 0
    for k, v in g.items():
        if k.startswith("bn/") and ("last_12" not in k and "last_24" not in k):
            ref = torch.from_numpy(v)
            got = sd[k[3:]]
            tol = 1e-4 * max(1.0, float(ref.abs().max()))
            assert float((got - ref).abs().max()) <= tol, k
            n += 1
    assert n > 150


def test_disp_head_independent_numpy_statement():
    rng = np.random.default_rng(3)
    mat = rng.standard_normal((2, 1, 8, 5, 7)).astype(np.float32) * 3
    for maxdisp in (24, 25):
        a = O.disp_head(torch.from_numpy(mat), maxdisp).numpy()
        b = O.disp_head_numpy_small(mat, maxdisp)
        assert np.abs(a - b).max() < 1e-4


def test_cost_volume_edges():
    x = np.arange(2 * 3 * 2 * 5, dtype=np.float32).reshape(2, 3, 2, 5) + 1
    y = -x
    c = O.cost_volume_numpy(x, y, 9)          # D3 = 3
    assert c.shape == (2, 6, 3, 2, 5)
    assert np.all(c[:, :, 2, :, :2] == 0)
    assert np.array_equal(c[:, :3, 2, :, 2:], x[:, :, :, 2:])
    assert np.array_equal(c[:, 3:, 2, :, 2:], y[:, :, :, :3])
    wide = O.cost_volume_numpy(x, y, 30)      # D3 = 10 > W: planes d >= W stay zero
    assert np.all(wide[:, :, 5:] == 0)
    assert O.cost_volume_numpy(x, y, 2).shape[2] == 0   # int(2/3) == 0 -> empty volume


def test_io_oracle_metrics_against_reference_fixtures():
    """oracle/io_oracle.py against numbers produced by running utils/metrics.py and evaluation.py:290-292
    (tests/golden/io_metrics.npz, written by tests/golden/make_io_golden.py)."""
    from oracle import io_oracle as IO
    z = load_golden("io_metrics")
    assert int(z["n_cases"]) >= 3
    for i in range(int(z["n_cases"])):
        p, t, maxdisp = z["pred%d" % i], z["true%d" % i], int(z["maxdisp%d" % i])
        want = z["result%d" % i]
        assert IO.three_px_error(p, t, maxdisp) == want[0]
        for k, thr in enumerate((1, 2, 3, 5)):
            assert IO.bad_pixel_frac(p, t, maxdisp, thr) == want[1 + k]
        assert IO.epe(p, t, maxdisp) == want[5]
        assert float(IO.validity_mask(t, maxdisp).sum()) == want[6]


def test_io_oracle_predict_against_reference_fixtures():
    """normalize_pair / test_transform against outputs of predict.py's own load_data / test_transform source."""
    from oracle import io_oracle as IO
    z = load_golden("io_predict")
    for i in range(int(z["n_cases"])):
        data = IO.normalize_pair(z["left%d" % i], z["right%d" % i])
        assert data.dtype == np.float32 and np.array_equal(data, z["data%d" % i])
        ch, cw = (int(v) for v in z["crop%d" % i])
        left, right = IO.test_transform(data, ch, cw)
        assert np.array_equal(left, z["input1_%d" % i]) and np.array_equal(right, z["input2_%d" % i])


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DIR), reason="reference tree only exists in the build container")
def test_io_oracle_against_live_reference():
    """Same two comparisons on fresh random data against the LIVE reference functions (utils/metrics.py imported,
    predict.py's two functions and evaluation.py's EPE lines executed from their own source)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_io_golden as G
    from oracle import io_oracle as IO
    M = G.reference_metrics()
    rs = np.random.RandomState(123)
    for maxdisp in (48, 192):
        t = (rs.rand(3, 40, 50) * maxdisp * 1.1).astype(np.float32)
        t[rs.rand(*t.shape) < 0.2] = 0
        p = (t + rs.randn(*t.shape) * 1.7).astype(np.float32)
        assert IO.three_px_error(p, t, maxdisp) == M.calculate_3px_error(p, t, maxdisp)
        assert IO.three_px_error(p, t, maxdisp) == M.calculate_3px_error_and_correct_mask(p, t, maxdisp)[0]
        for thr in (1, 2, 3):
            assert IO.bad_pixel_frac(p, t, maxdisp, thr) == M.calculate_bad_pixel_frac(p, t, maxdisp, thr)
        assert np.array_equal(IO.validity_mask(t, maxdisp), M.calculate_validity_mask(t, maxdisp))
        assert IO.epe(p, t, maxdisp) == G.reference_epe(p, t, maxdisp)
    left = rs.randint(0, 256, size=(19, 27, 3)).astype(np.uint8)
    right = rs.randint(0, 256, size=(19, 27, 3)).astype(np.uint8)
    test_transform, load_data = G.reference_predict_functions({"L": left, "R": right})
    data = load_data("L", "R")
    assert np.array_equal(IO.normalize_pair(left, right), data)
    for ch, cw in ((24, 30), (12, 18)):
        in1, in2, _, _ = test_transform(data, ch, cw)
        l, r = IO.test_transform(data, ch, cw)
        assert np.array_equal(l, in1.numpy()) and np.array_equal(r, in2.numpy())


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DIR), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = make_golden.build_reference(48).eval()
    left, right = make_golden.make_inputs(1, 48, 96)
    g = torch.Generator().manual_seed(7)
    left = left + 0.1 * torch.randn(left.shape, generator=g)
    with torch.no_grad():
        ref = model(left, right)
    disp = O.leastereo_forward(model.state_dict(), left, right, 48)
    assert float((disp - ref).abs().max()) <= 1e-3
