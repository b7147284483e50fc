"""Development aid: what bounds the single 16-/8-channel 3x3x3 convs (folded vs term-by-term weight images are chosen by
LEA_TC_FOLD in the environment; epilogue ablations through lea_tc_opts.debug - bit 0 no stores, bit 1 no TMEM loads, bit 2 no
residual reads, bit 3 no MMAs - only in a library built with -DLEA_TC_ABLATION)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import get_ops, PlanesVol, lea_tc_opts  # noqa: E402
from quick_perf import timeit  # noqa: E402

ops = get_ops()
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "4"))
print("LEA_TC_FOLD =", os.environ.get("LEA_TC_FOLD", "1"))
cases = [("L1op", 16, 16, (32, 64, 208)), ("L0op", 8, 8, (64, 128, 416)), ("L0x3", 8, 24, (64, 128, 416))]
for name, ci, co, sp in cases:
    src = PlanesVol.empty(B, ci, 2, *sp, dev)
    src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
    w = torch.randn(co, ci, 3, 3, 3, device=dev) * 0.05
    dst = PlanesVol.empty(B, co, 2, *sp, dev)
    dst.t.zero_()
    sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
    img = ops.pack_weights_tc(w, 2)
    for res in (True, False):
        p = ops.make_conv(src, 0, ci, co, 3, sc, sh, True, dst=dst, res=dst if res else None)
        for knobs in [{}, {"debug": 1}, {"debug": 3}, {"debug": 4}, {"debug": 7}, {"debug": 8}, {"debug": 15}, {"mma_terms": 1},
                      {"mma_terms": 1, "debug": 7}]:
            opts = lea_tc_opts()
            for kk, v in knobs.items():
                setattr(opts, kk, v)
            try:
                ms = timeit(lambda: ops.conv3d_tc(p, img, opts, w), iters=5, warm=1)
                print("%-5s res=%d %-40s %7.1f us" % (name, res, knobs, ms * 1e3), flush=True)
            except Exception as e:  # noqa: BLE001
                print(name, res, knobs, "failed:", str(e)[:100], flush=True)
    del src, dst
