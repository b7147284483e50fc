"""Generate golden input/output vectors for the hot path by running the UNMODIFIED reference
(/root/reference, devmentality/LEAStereo) on CPU in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference cannot travel to the GPU box, so the vectors are committed.  Weights are NOT stored (7 MB): they
are re-created from ``torch.manual_seed(0)`` by building the model, and pinned by a SHA-256 of the state_dict
stored in the fixture.  Per case we store: inputs, both feature maps, the matching cost ``mat``, the disparity,
a SHA-256 + strided sample of the cost volume, and (calibrated regime) every BN running_mean / running_var.
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = os.environ.get("LEASTEREO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, H, W, maxdisp, regime
    ("raw_48x96_d48", 48, 96, 48, "raw"),
    ("cal_48x96_d48", 48, 96, 48, "calibrated"),
    ("cal_46x94_d50", 46, 94, 50, "calibrated"),     # ragged: H, W, maxdisp not multiples of 3 (H3=16, W3=32, D3=16)
    ("cal_b2_24x48_d24", 24, 48, 24, "calibrated"),  # batch 2, smallest volume (level-2 depth 2)
    ("cal_skip_b2_24x48_d24", 24, 48, 24, "calibrated"),   # a 3D genotype with skip_connect ops (SKIP_GENOTYPE_3D)
]

# models/genotypes_3d.py:5-8: op 0 = skip_connect (Identity), op 1 = 3d_conv_3x3.  The shipped matching genotype has no
# op 0; this one exercises Identity inside a step sum (rows 0 and 3) and - through the row-order / branch-order quirk
# of skip_model_3d.py:33-36 vs :58-68 - an Identity that is the first contribution of a state as well as one that is
# added to a conv's output.
SKIP_GENOTYPE_3D = np.array([[1, 0], [0, 1], [3, 1], [4, 0], [8, 1], [6, 1]], dtype=np.int64)


def state_dict_sha256(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_reference(maxdisp, matching_genotype=None):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from config_utils.leastereo_args import LEAStereoArgs
    from retrain.LEAStereo import LEAStereo
    A = os.path.join(REF, "run/sceneflow/best/architecture/")
    cell_arch_mat = A + "matching_genotype.npy"
    if matching_genotype is not None:
        import tempfile
        cell_arch_mat = os.path.join(tempfile.mkdtemp(prefix="lea_genotype_"), "matching_genotype.npy")
        np.save(cell_arch_mat, np.asarray(matching_genotype))
    args = LEAStereoArgs(net_arch_fea=A + "feature_network_path.npy", cell_arch_fea=A + "feature_genotype.npy",
                         net_arch_mat=A + "matching_network_path.npy", cell_arch_mat=cell_arch_mat)
    args.maxdisp = maxdisp
    args.cuda = False
    torch.manual_seed(0)
    return LEAStereo(args, "cpu")


def make_inputs(B, H, W):
    g = torch.Generator().manual_seed(1)
    return torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)


def main():
    torch.set_num_threads(8)
    want = set(sys.argv[1:])
    for name, H, W, maxdisp, regime in CASES:
        if want and name not in want:
            continue
        B = 2 if "_b2_" in name else 1
        genotype = SKIP_GENOTYPE_3D if "_skip_" in name else None
        model = build_reference(maxdisp, genotype)
        sha_init = state_dict_sha256(model.state_dict())
        left, right = make_inputs(B, H, W)
        bn = {}
        if regime == "calibrated":
            # one fp32 train-mode pass with BN momentum 1.0, then eval
            for m in model.modules():
                if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                    m.momentum = 1.0
            model.train()
            with torch.no_grad():
                model(left, right)
            for k, v in model.state_dict().items():
                if k.endswith("running_mean") or k.endswith("running_var"):
                    bn[k] = v.numpy().copy()
        model.eval()
        stages = {}
        hooks = [model.matching.register_forward_hook(lambda m, i, o: stages.update(cost=i[0], mat=o))]
        with torch.no_grad():
            fx = model.feature(left)
            fy = model.feature(right)
            disp = model(left, right)
        for h in hooks:
            h.remove()
        cost = stages["cost"].numpy()
        out = dict(
            left=left.numpy(), right=right.numpy(), fx=fx.numpy(), fy=fy.numpy(), mat=stages["mat"].numpy(),
            disp=disp.numpy(), maxdisp=np.int64(maxdisp), regime=np.array(regime),
            cost_sha256=np.array(hashlib.sha256(cost.tobytes()).hexdigest()),
            cost_shape=np.array(cost.shape), cost_sample=cost[:, ::7, ::3, ::5, ::3].copy(),
            state_sha256_init=np.array(sha_init), torch_version=np.array(torch.__version__),
        )
        if genotype is not None:
            out["matching_genotype"] = genotype
        for k, v in bn.items():
            out["bn/" + k] = v
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "disp mean %.4f std %.4f" % (disp.mean(), disp.std()), "mat std %.4g" % stages["mat"].std(),
              os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
