"""Development aid: per-item cost of the tensor-core conv kernel's pipeline with no work (debug 11), item geometry varied."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import Ops, PlanesVol, lea_tc_opts  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = Ops(os.path.join(ROOT, "leastereo_b200", "_C", "libleastereo_b200_abl.so"))
dev = torch.device("cuda:0")
P, ci, co = 2, 8, 8
for k in (3, 1):
    for N, D, H, W, knobs in ((16, 1, 128, 416, {}), (1, 1, 2048, 416, {}), (1, 16, 128, 416, {}), (1, 16, 128, 416, {"depth_chunk": 1}),
                              (1, 16, 128, 416, {"depth_chunk": 2}), (1, 16, 128, 416, {"depth_chunk": 4}),
                              (16, 1, 128, 416, {"num_sms": 74}), (16, 1, 128, 416, {"num_sms": 37})):
        w = torch.zeros(co, ci, k, k, k, device=dev)
        w[:, :, k // 2] = torch.randn(co, ci, k, k, device=dev) * 0.05
        sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
        img = ops.pack_weights_tc(w, P)
        src = PlanesVol.empty(N, ci, P, D, H, W, dev)
        src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
        dst = PlanesVol.empty(N, co, P, D, H, W, dev)
        dst.t.zero_()
        p = ops.make_conv(src, 0, ci, co, k, sc, sh, True, dst=dst)
        row = []
        for dbg in (0, 11):
            opts = lea_tc_opts(); opts.debug = dbg
            for kk, v in knobs.items():
                setattr(opts, kk, v)
            row.append(timeit(lambda: ops.conv3d_tc(p, img, opts, w), iters=10, warm=2) * 1e3)
        print("k=%d N=%2d D=%2d H=%4d %-22s full %7.1f  ring-only %7.1f us" % (k, N, D, H, knobs, row[0], row[1]), flush=True)
        del src, dst
