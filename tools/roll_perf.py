"""Development aid: rolling vs chunked schedule of the tcgen05 conv at KITTI shapes, vs depth-chunk length."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import get_ops, PlanesVol, lea_tc_opts  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = get_ops()
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "4"))
cases = [("stem1", 32, 32, (64, 128, 416), False), ("L1op", 16, 16, (32, 64, 208), True), ("L0op", 8, 8, (64, 128, 416), True),
         ("L2op", 32, 32, (16, 32, 104), True), ("L1x3", 16, 48, (32, 64, 208), False), ("conv1", 128, 64, (32, 64, 208), False)]
for name, ci, co, sp, res in cases:
    src = PlanesVol.empty(B, ci, 2, *sp, dev)
    src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
    w = torch.randn(co, ci, 3, 3, 3, device=dev) * 0.05
    dst = PlanesVol.empty(B, co, 2, *sp, dev)
    dst.t.zero_()
    sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
    p = ops.make_conv(src, 0, ci, co, 3, sc, sh, True, dst=dst, res=dst if res else None)
    img = ops.pack_weights_tc(w, 2)
    for knobs in [{"early_drain": 2}, {}]:
        opts = lea_tc_opts()
        for kk, v in knobs.items():
            setattr(opts, kk, v)
        ms = timeit(lambda: ops.conv3d_tc(p, img, opts, w), iters=5, warm=1)
        print(name, B, knobs, "%.1f us/pair" % (ms * 1e3 / B), flush=True)
    del src, dst
