"""Development aid: fixed per-launch cost vs per-item cost of depth-1 convs (ablation library)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import Ops, PlanesVol, lea_tc_opts  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = Ops(os.path.join(ROOT, "leastereo_b200", "_C", "libleastereo_b200_abl.so"))
dev = torch.device("cuda:0")
for P in (2, 3):
    for (ci, co, k) in ((8, 8, 3), (16, 16, 3)):
        for N, D, H in ((16, 1, 16), (16, 1, 64), (16, 1, 128), (16, 1, 256), (1, 16, 128), (1, 64, 128)):
            W = 416
            w = torch.zeros(co, ci, k, k, k, device=dev)
            w[:, :, k // 2] = torch.randn(co, ci, k, k, device=dev) * 0.05
            sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
            img = ops.pack_weights_tc(w, P)
            src = PlanesVol.empty(N, ci, P, D, H, W, dev)
            src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
            dst = PlanesVol.empty(N, co, P, D, H, W, dev)
            dst.t.zero_()
            p = ops.make_conv(src, 0, ci, co, k, sc, sh, True, dst=dst)
            items = N * ((H + 15) // 16) * 52
            row = []
            for dbg in (0, 8, 11):
                opts = lea_tc_opts(); opts.debug = dbg
                row.append(timeit(lambda: ops.conv3d_tc(p, img, opts, w), iters=10, warm=2) * 1e3)
            print("P=%d %2d->%2d  N=%2d D=%2d H=%3d  tile-depths/SM %6.1f   full %7.1f  noMMA %7.1f  ring-only %7.1f us"
                  % (P, ci, co, N, D, H, items * D / 148.0, row[0], row[1], row[2]), flush=True)
            del src, dst
