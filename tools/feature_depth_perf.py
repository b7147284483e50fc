"""Development aid: a 2-D 3x3 conv over N images as N depth-1 volumes (what FeaturePlan runs) against the same images
stacked along the DEPTH axis of one volume (kd != 1 taps zero), 3 planes, tensor-core kernel."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import get_ops, PlanesVol, lea_tc_opts  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = get_ops()
dev = torch.device("cuda:0")
N = int(os.environ.get("N", "16"))
P = int(os.environ.get("P", "3"))
cases = [("8->8 1/3", 8, 8, 3, (128, 416)), ("16->16 1/6", 16, 16, 3, (64, 208)), ("32->32 1/3", 32, 32, 3, (128, 416)),
         ("24->8 1x1 1/3", 32, 8, 1, (128, 416)), ("48->16 1x1 1/6", 48, 16, 1, (64, 208))]
for name, ci, co, k, (H, W) in cases:
    w = torch.zeros(co, ci, k, k, k, device=dev)
    w[:, :, k // 2] = torch.randn(co, ci, k, k, device=dev) * 0.05
    sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
    img = ops.pack_weights_tc(w, P)
    for layout, (B, D) in (("batch", (N, 1)), ("depth", (1, N))):
        src = PlanesVol.empty(B, ci, P, D, H, W, dev)
        src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
        dst = PlanesVol.empty(B, co, P, D, H, W, dev)
        dst.t.zero_()
        for res in (False, True):
            p = ops.make_conv(src, 0, ci, co, k, sc, sh, True, dst=dst, res=dst if res else None)
            ms = timeit(lambda: ops.conv3d_tc(p, img, lea_tc_opts(), w), iters=10, warm=2)
            print("%-16s %-6s res=%d  %7.1f us" % (name, layout, res, ms * 1e3), flush=True)
        del src, dst
