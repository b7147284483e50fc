"""Development aid: full-forward tolerance vs the golden vectors for feature-net plane counts."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_checks as K
from conftest import GOLDEN_CASES, load_golden
from oracle import leastereo_oracle as O
DEV = torch.device("cuda:0")
for fp in (3, 2):
    for name in GOLDEN_CASES:
        g = load_golden(name)
        model = K.seeded_model(int(g["maxdisp"]))
        model.load_state_dict(K.golden_state_dict(g, model))
        model = model.to(DEV).eval()
        model.engine_options = {"feature_planes": fp}
        with torch.no_grad():
            d = model(torch.from_numpy(g["left"]).to(DEV), torch.from_numpy(g["right"]).to(DEV))
        rep = O.tolerance_report(d.cpu(), torch.from_numpy(g["disp"]))
        print("feature_planes", fp, name, {k: (round(v, 5) if isinstance(v, float) else v) for k, v in rep.items()})
