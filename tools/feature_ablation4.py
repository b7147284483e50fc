"""Development aid: per-role loop / barrier-wait cycles of CTA 0 of the tensor-core conv (ablation library)."""
import sys, os, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from leastereo_b200.kernels import Ops, PlanesVol, lea_tc_opts  # noqa: E402

ops = Ops(os.path.join(ROOT, "leastereo_b200", "_C", os.environ.get("ABL_LIB", "libleastereo_b200_abl.so")))
dev = torch.device("cuda:0")
cases = [("feat 8->8 3x3 P3", 16, 8, 8, 3, 3, (1, 128, 416)), ("feat 32->8 1x1 P3", 16, 32, 8, 1, 3, (1, 128, 416)),
         ("L1 16->16 P2", 4, 16, 16, 3, 2, (32, 64, 208)), ("L0 8->8 P2", 4, 8, 8, 3, 2, (64, 128, 416)),
         ("L1 64->16 1x1 P2", 4, 64, 16, 1, 2, (32, 64, 208)), ("stem1 32->32 P2", 2, 32, 32, 3, 2, (64, 128, 416)), ("conv1 128->64 P2", 4, 128, 64, 3, 2, (32, 64, 208)),
         ("stem0 64->32 P2", 1, 64, 32, 3, 2, (64, 128, 416))]
for name, N, ci, co, k, P, (D, H, W) in cases:
    w = torch.randn(co, ci, k, k, k, device=dev) * 0.05
    sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
    img = ops.pack_weights_tc(w, P)
    src = PlanesVol.empty(N, ci, P, D, H, W, dev)
    src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
    dst = PlanesVol.empty(N, co, P, D, H, W, dev)
    dst.t.zero_()
    res = os.environ.get("RES", "0") == "1"
    p = ops.make_conv(src, 0, ci, co, k, sc, sh, True, dst=dst, res=dst if res else None)
    for dbg in (0, 11):
        opts = lea_tc_opts(); opts.debug = dbg
        for _ in range(3):
            ops.conv3d_tc(p, img, opts, w)
        torch.cuda.synchronize()
        out = (ctypes.c_longlong * 12)()
        ops.lib.lea_tc_prof(out)
        r = list(out)
        def fmt(i):
            tot, wait, items = r[3 * i:3 * i + 3]
            return "%8d cyc, wait %5.1f %%, %4d items, busy/item %6.0f" % (tot, 100.0 * wait / max(tot, 1), items, (tot - wait) / max(items, 1))
        print("%-20s dbg=%2d  producer[%s]  issuer[%s]  epilogue[%s]" % (name, dbg, fmt(0), fmt(1), fmt(2)), flush=True)
    del src, dst
