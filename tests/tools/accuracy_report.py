"""Tolerance table of the product hot path against the golden vectors (reference outputs), per golden case and engine
mode.  North-star tolerance: |dd| <= 0.1 px on >= 99.9 % of pixels and mean |dd| <= 0.01 px."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_checks as K  # noqa: E402
from conftest import GOLDEN_CASES  # noqa: E402
from leastereo_b200.kernels import get_ops  # noqa: E402


def main():
    ops = get_ops()
    dev = torch.device("cuda:0")
    modes = [("simt", 3, 0), ("simt", 2, 0), ("tc", 3, 0), ("tc", 2, 0), ("tc", 2, 1)]
    extra = None
    args = sys.argv[1:]
    if "--extra" in args:                       # engine-option overrides as JSON, e.g. --extra '{"accum_split": 2}'
        i = args.index("--extra")
        extra = json.loads(args[i + 1])
        args = args[:i] + args[i + 2:]
    if args:
        modes = [m for m in modes if m[0] in args]
    out = []
    print("%-20s %-6s %-2s %-5s | %-9s %-9s %-8s %-9s %s" % ("case", "conv", "P", "terms", "frac<=0.1", "mean", "max", "mat_rel", "ok"))
    for name in GOLDEN_CASES:
        for conv, planes, terms in modes:
            rep = K.check_hot_path_golden(ops, dev, name, planes=planes, conv=conv, mma_terms=terms, mat_rtol=None,
                                          require_tolerance=False, extra=extra)
            print("%-20s %-6s %-2d %-5d | %-9.5f %-9.5f %-8.4f %-9.2e %s" % (
                name, conv, planes, terms, rep["frac_within_0p1"], rep["mean_abs"], rep["max_abs"], rep["mat_rel_err"],
                "PASS" if rep["ok"] else "FAIL"))
            out.append(dict(case=name, conv=conv, planes=planes, mma_terms=terms, **rep))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "accuracy_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
