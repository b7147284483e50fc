"""Development aid: accumulator sets x depth chunk of the big tensor-core convs (KITTI shapes, 8 pairs)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import get_ops, PlanesVol, lea_tc_opts  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = get_ops()
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "8"))
cases = [("conv1 128->64 L1", 128, 64, (32, 64, 208), [(0, 0), (1, 4), (2, 2), (1, 2), (2, 1), (1, 3)]),
         ("stem1 32->32 L0", 32, 32, (64, 128, 416), [(0, 0), (2, 4), (1, 8), (2, 2), (1, 4), (2, 3), (1, 6)]),
         ("batched 16->48 L1", 16, 48, (32, 64, 208), [(0, 0), (2, 2), (1, 5), (1, 4), (1, 3), (2, 1)]),
         ("L2 32->32", 32, 32, (16, 32, 104), [(0, 0), (2, 4), (1, 8), (2, 2), (1, 4)]),
         ("L1 16->16", 16, 16, (32, 64, 208), [(0, 0), (2, 8), (2, 4), (1, 16), (1, 8)])]
for name, ci, co, sp, knobs in cases:
    src = PlanesVol.empty(B, ci, 2, *sp, dev)
    src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
    w = torch.randn(co, ci, 3, 3, 3, device=dev) * 0.05
    dst = PlanesVol.empty(B, co, 2, *sp, dev)
    sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
    img = ops.pack_weights_tc(w, 2)
    p = ops.make_conv(src, 0, ci, co, 3, sc, sh, True, dst=dst)
    for sets, dc in knobs:
        opts = lea_tc_opts(); opts.acc_sets = sets; opts.depth_chunk = dc
        try:
            ms = timeit(lambda: ops.conv3d_tc(p, img, opts, w), iters=6, warm=2)
            print("%-20s acc_sets=%d depth_chunk=%d  %8.1f us" % (name, sets, dc, ms * 1e3), flush=True)
        except Exception as e:  # noqa: BLE001
            print(name, sets, dc, "failed:", str(e)[:80], flush=True)
    del src, dst
