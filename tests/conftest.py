import hashlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["raw_48x96_d48", "cal_48x96_d48", "cal_46x94_d50", "cal_b2_24x48_d24", "cal_skip_b2_24x48_d24"]
REFERENCE_DIR = "/root/reference"   # exists only in the build container; never read by -m gpu tests


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def state_dict_sha256(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def seeded_model(maxdisp, device="cpu", matching_genotype=None):
    """The product module with the same seeded random init as the reference (torch.manual_seed(0));
    ``matching_genotype`` replaces the shipped 3D cell genotype (models/genotypes_3d.py:5-8 op indices)."""
    import contextlib, io, tempfile
    from leastereo_b200 import LEAStereo, default_args
    args = default_args(maxdisp=maxdisp, cuda=(device != "cpu"))
    if matching_genotype is not None:
        path = os.path.join(tempfile.mkdtemp(prefix="lea_genotype_"), "matching_genotype.npy")
        np.save(path, np.asarray(matching_genotype))
        args.cell_arch_mat = path
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = LEAStereo(args, device)
    return model


def golden_genotype(g):
    """The 3D cell genotype a golden case was generated with (None = the shipped one)."""
    return g["matching_genotype"] if "matching_genotype" in g else None


def golden_arch(g):
    """Architecture dict for the oracle's functional forward (oracle.leastereo_oracle.SHIPPED_ARCH layout)."""
    from oracle import leastereo_oracle as O
    arch = dict(O.SHIPPED_ARCH)
    if "matching_genotype" in g:
        arch["matching_genotype"] = np.asarray(g["matching_genotype"])
    return arch


def golden_model(g, device="cpu"):
    """Seeded product module for a golden case (its genotype), fixture state loaded."""
    model = seeded_model(int(g["maxdisp"]), matching_genotype=golden_genotype(g))
    model.load_state_dict(golden_state_dict(g, model))
    return model.to(device)


def golden_state_dict(g, model=None):
    """state_dict for a golden case: seeded init (+ the fixture's calibrated BN statistics)."""
    model = model or seeded_model(int(g["maxdisp"]), matching_genotype=golden_genotype(g))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    assert state_dict_sha256(sd) == str(g["state_sha256_init"]), "seeded init differs from the reference's"
    for k, v in g.items():
        if k.startswith("bn/"):
            sd[k[3:]] = torch.from_numpy(v.copy())
    return sd


@pytest.fixture(scope="session")
def golden():
    return {n: load_golden(n) for n in GOLDEN_CASES}
