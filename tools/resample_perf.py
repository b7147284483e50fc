"""Development aid: up-sampling resamples of the KITTI plan (conv-before-upsample: raw low-res conv output -> BN+ReLU)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import get_ops, PlanesVol  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = get_ops()
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "8"))
for name, c, s_in, s_out in (("8ch L1->L0 (cell 10)", 8, (32, 64, 208), (64, 128, 416)), ("16ch L2->L1 (cells 4,8,9)", 16, (16, 32, 104), (32, 64, 208)),
                            ("32ch L0->L1 (down)", 32, (64, 128, 416), (32, 64, 208))):
    src = PlanesVol.empty(B, c, 2, *s_in, dev)
    src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16())
    dst = PlanesVol.empty(B, c, 2, *s_out, dev)
    sc = torch.ones(c, device=dev); sh = torch.zeros(c, device=dev)
    ms = timeit(lambda: ops.trilinear_ac(src, 0, c, dst, 0, sc, sh, True), iters=10, warm=2)
    nbytes = 4.0 * B * c * (s_in[0] * s_in[1] * s_in[2] + s_out[0] * s_out[1] * s_out[2])
    print("%-28s %7.1f us  %6.0f GB/s" % (name, ms * 1e3, nbytes / ms / 1e6), flush=True)
    del src, dst
