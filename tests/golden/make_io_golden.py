"""Golden vectors for the callers on either side of the hot path, produced by RUNNING THE
REFERENCE'S OWN CODE in the build container:

    python tests/golden/make_io_golden.py         # writes tests/golden/io_metrics.npz, tests/golden/io_predict.npz

* ``utils/metrics.py`` imports cleanly (numpy only) and is called as is.
* ``predict.py`` cannot be imported (skimage, matplotlib, path are absent and it parses argv at import), so the source
  lines of its two functions ``test_transform`` (:144-159) and ``load_data`` (:162-184) are cut out of the file by
  ``ast`` and executed unmodified in a namespace whose ``Image.open`` returns in-memory arrays instead of reading files.
* ``evaluation.py:290-292`` (the EPE expression) is likewise executed from its own source lines.

Nothing of the reference is copied into the repository: the script reads the files where they lie under /root/reference
at generation time, and only inputs / outputs are stored.
"""
import ast
import os
import sys

import numpy as np
import torch

REF = os.environ.get("LEASTEREO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_metrics():
    sys.path.insert(0, REF)
    import importlib
    return importlib.import_module("utils.metrics")


def reference_predict_functions(images):
    """predict.py's test_transform / load_data, executed from their own source; ``images`` maps file name -> uint8 HWC."""
    path = os.path.join(REF, "predict.py")
    src = open(path).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("test_transform", "load_data")]
    assert len(wanted) == 2
    mod = ast.Module(body=wanted, type_ignores=[])

    class _Image:
        @staticmethod
        def open(name):
            return images[name]

    ns = {"np": np, "torch": torch, "Image": _Image}
    exec(compile(mod, path, "exec"), ns)
    return ns["test_transform"], ns["load_data"]


def reference_epe(prediction, disp, maxdisp):
    """evaluation.py:290-291 executed from its own two source lines."""
    path = os.path.join(REF, "evaluation.py")
    lines = open(path).read().splitlines()
    sel = [ln.strip() for ln in lines if ln.strip().startswith("mask = np.logical_and(disp >= 0.001")
           or ln.strip().startswith("error = np.mean(np.abs(prediction[mask] - disp[mask]))")]
    assert len(sel) == 2, sel

    class _Opt:
        pass
    opt = _Opt()
    opt.maxdisp = maxdisp
    ns = {"np": np, "prediction": prediction, "disp": disp, "opt": opt}
    exec("\n".join(sel), ns)
    return float(ns["error"])


def metric_cases():
    """(pred, true, maxdisp) cases: errors on both sides of every integer threshold, targets at the mask boundaries
    (0, 0.001, maxdisp), large targets whose 5 % band exceeds 3 px, targets so large that the reference counts the
    INVALID pixel as correct (10000 < 0.05 t), float64 and float32 inputs."""
    g = np.random.RandomState(11)
    cases = []
    for shape, maxdisp, dtype in [((37, 53), 192, np.float32), ((2, 24, 31), 96, np.float32), ((16, 16), 408, np.float32)]:
        t = (g.rand(*shape) * maxdisp * 1.2).astype(dtype)
        t[g.rand(*shape) < 0.15] = 0.0
        t.flat[:6] = [0.001, 0.0010001, maxdisp, maxdisp - 1e-3, 250000.0, 199999.0]
        p = (t + g.randn(*shape) * 2.0).astype(dtype)
        p.flat[10:20] = t.flat[10:20] + np.array([0.999, 1.0, 1.001, 1.999, 2.0, 2.999, 3.0, 3.001, -1.0, -2.5], dtype)
        p[g.rand(*shape) < 0.05] += 9.0
        cases.append((p, t, maxdisp))
    return cases


def main():
    M = reference_metrics()
    out = {}
    for i, (p, t, maxdisp) in enumerate(metric_cases()):
        out["pred%d" % i], out["true%d" % i], out["maxdisp%d" % i] = p, t, np.int64(maxdisp)
        res = [M.calculate_3px_error(p, t, maxdisp)]
        e2, correct = M.calculate_3px_error_and_correct_mask(p, t, maxdisp)
        assert e2 == res[0]
        res += [M.calculate_bad_pixel_frac(p, t, maxdisp, thr) for thr in (1, 2, 3, 5)]
        res.append(reference_epe(p, t, maxdisp))
        res.append(float(M.calculate_validity_mask(t, maxdisp).sum()))
        out["result%d" % i] = np.array(res, dtype=np.float64)     # [3px, bad1, bad2, bad3, bad5, epe, #valid]
        out["correct%d" % i] = np.packbits(correct.reshape(-1))
        print("metrics case", i, res)
    out["n_cases"] = np.int64(len(metric_cases()))
    np.savez_compressed(os.path.join(HERE, "io_metrics.npz"), **out)

    g = np.random.RandomState(7)
    out = {}
    shapes = [(20, 30, 24, 36), (20, 30, 20, 30), (20, 30, 16, 24), (37, 53, 24, 48), (5, 300, 8, 512), (21, 33, 21, 30)]
    for i, (h, w, ch, cw) in enumerate(shapes):
        left = g.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        right = g.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        left[:, :, 1] //= 4
        test_transform, load_data = reference_predict_functions({"L": left, "R": right})
        data = load_data("L", "R")
        in1, in2, hh, ww = test_transform(data, ch, cw)
        assert (hh, ww) == (h, w)
        out["left%d" % i], out["right%d" % i] = left, right
        out["crop%d" % i] = np.array([ch, cw], dtype=np.int64)
        out["data%d" % i] = data
        out["input1_%d" % i], out["input2_%d" % i] = in1.numpy(), in2.numpy()
        print("predict case", i, (h, w, ch, cw), data.dtype, in1.shape)
    out["n_cases"] = np.int64(len(shapes))
    np.savez_compressed(os.path.join(HERE, "io_predict.npz"), **out)


if __name__ == "__main__":
    main()
