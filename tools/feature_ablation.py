"""Development aid: what bounds the depth-1 (2-D) convs of the feature net - lea_tc_opts.debug ablations (bit 0 no stores,
bit 1 no TMEM loads, bit 2 no residual reads, bit 3 no MMAs) on a library built with -DLEA_TC_ABLATION
(leastereo_b200/_C/libleastereo_b200_abl.so)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import Ops, PlanesVol, lea_tc_opts  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = Ops(os.path.join(ROOT, "leastereo_b200", "_C", "libleastereo_b200_abl.so"))
dev = torch.device("cuda:0")
N, P = 16, int(os.environ.get("P", "3"))
cases = [("8->8 3x3 1/3", 8, 8, 3, (128, 416)), ("32->8 1x1 1/3", 32, 8, 1, (128, 416)), ("16->16 3x3 1/6", 16, 16, 3, (64, 208))]
for name, ci, co, k, (H, W) in cases:
    w = torch.zeros(co, ci, k, k, k, device=dev)
    w[:, :, k // 2] = torch.randn(co, ci, k, k, device=dev) * 0.05
    sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
    img = ops.pack_weights_tc(w, P)
    src = PlanesVol.empty(N, ci, P, 1, H, W, dev)
    src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
    dst = PlanesVol.empty(N, co, P, 1, H, W, dev)
    dst.t.zero_()
    p = ops.make_conv(src, 0, ci, co, k, sc, sh, True, dst=dst)
    for knobs in [{}, {"debug": 1}, {"debug": 3}, {"debug": 8}, {"debug": 9}, {"debug": 11}, {"acc_sets": 1}, {"acc_sets": 2}]:
        opts = lea_tc_opts()
        for kk, v in knobs.items():
            setattr(opts, kk, v)
        try:
            ms = timeit(lambda: ops.conv3d_tc(p, img, opts, w), iters=10, warm=2)
            print("%-16s %-28s %7.1f us" % (name, knobs, ms * 1e3), flush=True)
        except Exception as e:  # noqa: BLE001
            print(name, knobs, "failed:", str(e)[:100], flush=True)
