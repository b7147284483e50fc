"""Development aid: tensor-core conv timings vs volume size (fixed launch cost vs streaming rate) and knobs."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import get_ops, PlanesVol, lea_tc_opts  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = get_ops()
dev = torch.device("cuda:0")
cases = [(16, 16, 3, (2, 16, 8)), (16, 16, 3, (8, 16, 8)), (16, 16, 3, (8, 64, 296)), (16, 16, 3, (16, 64, 296)),
         (16, 16, 3, (32, 64, 208)), (16, 16, 1, (1, 16, 8)), (16, 16, 1, (8, 64, 296)), (64, 16, 1, (8, 64, 296)),
         (32, 32, 3, (16, 32, 104)), (32, 32, 3, (8, 16, 8))]
for (ci, co, k, sp) in cases:
    src = PlanesVol.empty(1, ci, 2, *sp, dev)
    src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
    w = torch.randn(co, ci, k, k, k, device=dev) * 0.05
    dst = PlanesVol.empty(1, co, 2, *sp, dev)
    p = ops.make_conv(src, 0, ci, co, k, None, None, False, dst=dst)
    img = ops.pack_weights_tc(w, 2)
    for knobs in [{}, {"mma_terms": 1}]:
        opts = lea_tc_opts()
        for kk, v in knobs.items():
            setattr(opts, kk, v)
        ms = timeit(lambda: ops.conv3d_tc(p, img, opts, w), iters=9, warm=2)
        vox = sp[0] * sp[1] * sp[2]
        gb = vox * (ci + co) * 4 / 1e9
        print(ci, co, k, sp, knobs, "%.1f us  %.2f TB/s" % (ms * 1e3, gb / ms))
x = torch.zeros(1024, device=dev)
print("empty torch kernel: %.1f us" % (timeit(lambda: x.add_(1.0), iters=9) * 1e3))
