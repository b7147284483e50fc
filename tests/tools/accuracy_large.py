"""Tolerance table at BASELINE sizes: the product (images in, disparity out) against the reference-run fixtures
tests/golden/large_*.npz, per engine mode.  North-star tolerance: |dd| <= 0.1 px on >= 99.9 % of pixels, mean <= 0.01 px.

    python tests/tools/accuracy_large.py [case-substring ...] [--modes name,name] [--out file.json]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_state_dict, load_golden, seeded_model  # noqa: E402
from oracle import leastereo_oracle as O  # noqa: E402

CASES = ["large_cal_288x576_d192", "large_raw_288x576_d192", "large_cal_384x1248_d192"]
MODES = {
    "tc_p2_default": {},
    "tc_p2_plain_stem0": {"collapse_stem0": False},
    "tc_p2_no_head_taps": {"fuse_head": False},
    "tc_p2_one_region": {"accum_split": 2},
    "tc_p2_split3": {"accum_split": 3},
    "tc_p2_split4": {"accum_split": 4},
    "tc_p2_feature_p2": {"feature_planes": 2},
    "tc_p2_feature_terms2": {"feature_terms": 2},
    "tc_p3": {"planes": 3},
    "simt_p2": {"conv": "simt", "planes": 2},
    "simt_p3": {"conv": "simt", "planes": 3},
}


def pair_inputs(k, H, W):
    g = torch.Generator().manual_seed(1 + k)
    return torch.randn(1, 3, H, W, generator=g), torch.randn(1, 3, H, W, generator=g)


def main():
    args = sys.argv[1:]
    out_path = os.path.join(ROOT, "gpurun_out", "accuracy_large.json")
    modes = list(MODES)
    if "--modes" in args:
        i = args.index("--modes"); modes = args[i + 1].split(","); args = args[:i] + args[i + 2:]
    if "--out" in args:
        i = args.index("--out"); out_path = args[i + 1]; args = args[:i] + args[i + 2:]
    cases = [c for c in CASES if not args or any(a in c for a in args)]
    dev = torch.device("cuda:0")
    from leastereo_b200 import engine
    rows = []
    print("%-26s %-20s | %-9s %-9s %-8s %-9s %s" % ("case", "mode", "frac<=0.1", "mean", "max", "mat_rel", "ok"))
    for name in cases:
        g = load_golden(name)
        H, W, maxdisp = int(g["H"]), int(g["W"]), int(g["maxdisp"])
        left, right = pair_inputs(0, H, W)
        for mode in modes:
            model = seeded_model(maxdisp)
            model.load_state_dict(golden_state_dict(g, model))
            model = model.to(dev).eval()
            model.engine_options = dict(MODES[mode])
            with torch.no_grad():
                d = model(left.to(dev), right.to(dev))
            plan = next(p for p in engine._plans(model.matching).values())
            mat = plan.mat.cpu().numpy()[:, :, ::3, ::5, ::7]
            rep = O.tolerance_report(d.cpu(), torch.from_numpy(g["disp0"]))
            rep["mat_rel_err"] = float(np.abs(mat - g["mat_sample"]).max() / float(g["mat_absmax"]))
            rep["mat_rel_rms"] = float(np.sqrt(np.mean((mat.astype(np.float64) - g["mat_sample"]) ** 2)) /
                                       np.sqrt(np.mean(g["mat_sample"].astype(np.float64) ** 2)))
            print("%-26s %-20s | %-9.5f %-9.5f %-8.4f %-9.2e %s" % (name, mode, rep["frac_within_0p1"], rep["mean_abs"],
                                                                    rep["max_abs"], rep["mat_rel_err"],
                                                                    "PASS" if rep["ok"] else "FAIL"), flush=True)
            rows.append(dict(case=name, mode=mode, options=MODES[mode], **rep))
            del model, plan
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(rows, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
